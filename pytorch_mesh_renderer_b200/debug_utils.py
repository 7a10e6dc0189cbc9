"""Counterpart of the reference's src/common/debug_utils.py (two helpers used while developing the examples)."""
import torch


def debug_tensor(tensor, msg=""):
    """Prints a tensor in full (debug_utils.py:3-7)."""
    torch.set_printoptions(profile="full", linewidth=200)
    try:
        print("[debug tensor] {}".format(msg))
        print(tensor)
    finally:
        torch.set_printoptions(profile="default", linewidth=80)


def check_isnan_isinf(tensor, msg=""):
    """Raises ValueError(msg) if the tensor holds a NaN or an infinity (debug_utils.py:9-11); one device-side
    reduction and one read-back."""
    if not bool(torch.isfinite(tensor).all()):
        raise ValueError(msg)
