"""Synthetic workloads -- seeded checkpoints, crop batches, frames, detections and a stand-in SMPL model.
DATA ONLY: nothing here computes any part of the hot path; bench.py, smoke(), tools/ and tests/ share it."""
