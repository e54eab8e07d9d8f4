"""Host-side data path of the UNITER hot path: synthetic batches of the
BASELINE shapes, the fine-tuning collate (reference data/meme_dataset.py:152-214,
data/dataset_template.py:92-114) and a pinned, double-buffered host->device prefetcher."""
from .pipeline import PinnedPrefetcher, box7, collate_memes, load_img_feature  # noqa: F401,E402
from .synthetic import synth_batch, synth_pretrain_batch  # noqa: F401,E402
