"""Test infrastructure only: CPU restatement of the reference hot path (see sea_oracle.py)."""
