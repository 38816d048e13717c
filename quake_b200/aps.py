"""Adaptive Partition Scanning (APS) driver -- filled in after the fixed-nprobe path."""
from __future__ import annotations


def adaptive_scan(index, xq, p_ids, slots, sp):
    raise NotImplementedError("APS (recall_target > 0) is not implemented yet in quake_b200")
