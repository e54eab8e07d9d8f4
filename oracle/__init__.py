"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's UNITER hot path.

Nothing under oracle/ is imported by the product package (meme_challenge_b200/). Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the checker
or as the timed CPU baseline, never as the thing shipped.
"""
