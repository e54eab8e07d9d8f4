"""Generate tests/golden/metrics.npz by running the UNMODIFIED reference data/metrics.py (standard_metrics with
add_optimal_acc=True: accuracy / recall / precision / F1 / aucroc / optimal threshold) on seeded probability /
label vectors, including heavy ties. matplotlib / seaborn (plot-only imports of that module, not installed here)
are stubbed. Run in the build container only:  python oracle/make_golden_metrics.py"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("B200U_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "metrics.npz")


def main():
    for m in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.path.insert(0, REF)
    import data.metrics as M
    g = {}
    cases = []
    torch.manual_seed(0)
    p = torch.rand(200); cases.append((p, (p > torch.rand(200)).long()))
    p = torch.rand(64); cases.append((p, (torch.rand(64) < 0.36).long()))                      # uninformative scores
    p = (torch.rand(300) * 10).round() / 10; cases.append((p, (p + 0.3 * torch.randn(300) > 0.5).long()))  # many ties
    p = torch.sigmoid(torch.randn(1000) * 3); cases.append((p, (p > torch.rand(1000)).long()))
    p = torch.tensor([0.2, 0.8, 0.8, 0.1]); cases.append((p, torch.tensor([0, 1, 1, 0])))       # separable
    p = torch.tensor([0.9, 0.1, 0.5]); cases.append((p, torch.tensor([0, 1, 1])))               # anti-correlated
    for i, (p, l) in enumerate(cases):
        m = M.standard_metrics(p, l, add_optimal_acc=True)
        g["c%d_probs" % i] = p.numpy()
        g["c%d_labels" % i] = l.numpy()
        for k, v in m.items():
            g["c%d_%s" % (i, k)] = np.float64(v)
    g["n_cases"] = np.array(len(cases))
    np.savez_compressed(OUT, **g)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
