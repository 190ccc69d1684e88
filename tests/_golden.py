"""Loader for the fixtures minted from the reference by oracle/gen_golden.py (test helper)."""
import glob
import os

import numpy as np

from oracle import oracle

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.basename(p)[len("ldconv_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "ldconv_*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"ldconv_{name}.npz"))
    inc, outc, N, s, B, H, W = (int(v) for v in z["meta"])
    eps, momentum = (float(v) for v in z["bn_cfg"])
    prm = oracle.LDConvParams(
        p_conv_weight=z["param_p_conv_weight"].copy(), p_conv_bias=z["param_p_conv_bias"].copy(),
        conv_weight=z["param_conv_0_weight"].copy(), bn_weight=z["param_conv_1_weight"].copy(),
        bn_bias=z["param_conv_1_bias"].copy(), running_mean=z["param_conv_1_running_mean"].copy(),
        running_var=z["param_conv_1_running_var"].copy(), num_param=N, stride=s, eps=eps, momentum=momentum)
    meta = dict(inc=inc, outc=outc, N=N, s=s, B=B, H=H, W=W, h=(H - 1) // s + 1, w=(W - 1) // s + 1)
    return z, prm, meta
