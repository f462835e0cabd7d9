"""CPU oracle for the EMA row.  TEST INFRASTRUCTURE ONLY (same rules as
patchnce_oracle.py).  Restates utils/io_ckpt.py:9-53 of the reference:

* shadow[name] = param.clone() for every parameter that requires grad                      :19-21
* update():  shadow = (1 - decay) * param + decay * shadow, three fp32 roundings, no fma   :23-29
* apply_shadow() / restore(): swap the shadow values in and out of the model               :31-43

Pinned by tests/golden/ema_reference.npz, frozen from the unmodified reference by oracle/make_golden_ema.py.
"""
import numpy as np


def ema_update_np(shadow: np.ndarray, param: np.ndarray, decay: float) -> np.ndarray:
    """One update of one tensor in float32 arithmetic: fl(fl((1-decay)*p) + fl(decay*s))."""
    a = np.float32(1.0 - decay)          # torch turns the Python double into an fp32 scalar inside the kernel
    b = np.float32(decay)
    t1 = (a * param.astype(np.float32)).astype(np.float32)
    t2 = (b * shadow.astype(np.float32)).astype(np.float32)
    return (t1 + t2).astype(np.float32)
