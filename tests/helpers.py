"""Shared helpers for the parity tests."""
import numpy as np

TINY = 1e-300


def rel_err(got, ref, floor=1.0):
    """max |got-ref| / max(|ref|, floor*scale): component-wise relative error with an absolute floor so
    that components that are analytically zero (1e-16 noise) do not dominate.  floor is in units of the
    largest |ref| component of the same array."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    scale = max(float(np.max(np.abs(ref))) if ref.size else 0.0, TINY)
    den = np.maximum(np.abs(ref), floor * scale)
    return float(np.max(np.abs(got - ref) / den)) if ref.size else 0.0


def comp_rel_err(got, ref, atol):
    """strict per-component relative error: |got-ref| / max(|ref|, atol)"""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), atol))) if ref.size else 0.0
