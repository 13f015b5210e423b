"""Precision knob (the only setting the reference does not have).

precision = "auto": float32 inputs run the fp32 path (bf16x2-split tensor-core contractions, fp32
sweeps, fp64 scalar reductions); anything else runs the fp64 path, which reproduces the reference's
float64 results to ~1e-10.  "fp32" / "fp64" force a mode regardless of the input dtype.
"""
precision = "auto"


def set_precision(mode):
    global precision
    if mode not in ("auto", "fp32", "fp64"):
        raise ValueError(f"precision must be 'auto', 'fp32' or 'fp64', got {mode!r}")
    precision = mode
