"""experiment_yolo_b200 -- B200-native (sm_100a) LDConv for DEAL-YOLO (adityaX1412/Experiment-YOLO).

Only what the hot path needs: the CUDA kernels + C ABI (csrc/, include/ldconv_b200.h), the host-side
mirror of the reference's `LDConv` module and YAML hook (ldconv.py), the DEAL-YOLO-LD graph the benchmark runs
(dealyolo.py) and the batch-sharding / gradient all-reduce helpers (dist.py).
"""
from .ldconv import LDConv, install, ldconv_function  # noqa: F401

__all__ = ["LDConv", "install", "ldconv_function"]
