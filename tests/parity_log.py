"""Measured parity errors, kept where the driver and a reader can see them (VERDICT r1 'What's weak' #1).

Every GPU parity test calls record(...) / check_final_logits(...); the numbers are printed in pytest's terminal summary
(also with -q, no -s needed) and written as JSON to $RLCF_PARITY_LOG (default gpurun_out/r2_parity.json, which gpurun
merges back into the build container; the copy under profiles/r2_parity.json is the committed one).
"""
from __future__ import annotations

import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RECORDS: list = []


def _f(x):
    try:
        return float(x)
    except (TypeError, ValueError):
        return x


def record(test: str, **vals) -> dict:
    rec = {"test": test, **{k: (_f(v) if not isinstance(v, (str, bool, int, list, dict, type(None))) else v)
                            for k, v in vals.items()}}
    RECORDS.append(rec)
    return rec


def check_final_logits(tag, got, ref_final, scale, delta, tol=1e-3, allow_delta=0.0, why=None):
    """Direct comparison of adapted logits with the reference's: max|got - ref| <= tol * max|logit| (north_star's
    bound).  allow_delta > 0 adds that fraction of what adaptation changed (delta) to the bound -- only for the cases
    listed with a justification (`why`), which is recorded next to the measured numbers."""
    import numpy as np
    err = float(np.abs(np.asarray(got, dtype=np.float64) - np.asarray(ref_final, dtype=np.float64)).max())
    scale, delta = float(scale), float(delta)
    bound = tol * scale + allow_delta * delta
    record(tag, final_err_rel=err / scale, adaptation_delta_rel=delta / scale, tol=tol, allow_delta=allow_delta,
           strict_ok=bool(err <= tol * scale), why=why)
    assert err <= bound, (f"{tag}: adapted logits differ from the reference by {err / scale:.2e} of max|logit| "
                          f"(bound {bound / scale:.2e}; adaptation moved them by {delta / scale:.2e})")
    return err / scale


def dump(path: str | None = None) -> str | None:
    if not RECORDS:
        return None
    path = path or os.environ.get("RLCF_PARITY_LOG") or os.path.join(ROOT, "gpurun_out", "r2_parity.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump({"records": RECORDS}, f, indent=1)
    except OSError:
        return None
    return path


def summary_lines() -> list:
    out = []
    for r in RECORDS:
        kv = ", ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}" for k, v in r.items()
                       if k not in ("test", "why") and v is not None)
        out.append(f"{r['test']}: {kv}")
    return out
