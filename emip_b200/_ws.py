"""Caller-owned workspace helper: the C ABI never allocates device memory."""
import torch


def workspace(nbytes, device, align=1024):
    """(tensor keeping the memory alive, aligned data pointer, usable bytes)."""
    buf = torch.empty(int(nbytes) + align, dtype=torch.uint8, device=device)
    base = buf.data_ptr()
    off = (-base) % align
    return buf, base + off, int(nbytes)
