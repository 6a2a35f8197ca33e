"""Test infrastructure only: CPU restatement of the reference hot path (see oracle.py). Never imported by bitswap_b200/."""
