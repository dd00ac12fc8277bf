"""Driver entry points: build() compiles every CUDA translation unit for sm_100a (and loads the
library); smoke() runs one tiny pass of the hot path on cuda:0 and checks it against the oracle."""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build() -> None:
    from fullsubnet_b200.csrc.build import build as build_lib
    path = build_lib(force=bool(os.environ.get("FSN_FORCE_BUILD")))
    import fullsubnet_b200  # noqa: F401
    from fullsubnet_b200 import _lib
    lib = _lib.load()
    assert lib.fsn_built_arch() == 100, "library was not built for sm_100a"
    # the oracle is Python (torch CPU): "building" the checker = importing it
    from oracle import fullsubnet_oracle  # noqa: F401
    print(f"[build] {path} ok (abi v{lib.fsn_version()}, sm_{lib.fsn_built_arch()}a)")


def smoke() -> None:
    import numpy as np
    import torch

    from fullsubnet_b200.fullsubnet.model import Model
    from oracle import fullsubnet_oracle as O

    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    dev = torch.device("cuda:0")
    sd = O.make_state_dict(seed=0)
    y = O.make_noisy(2, 4000, seed=3, speechlike=True)
    ref_wav, ref_crm = O.enhance(y, sd, return_crm=True)
    for prec in ("fp32", "auto"):
        m = Model(**O.DEFAULT_MODEL_ARGS, precision=prec)
        m.load_state_dict(sd, strict=True)
        m = m.to(dev).eval()
        wav, crm = m.enhance(y.to(dev), return_crm=True)
        torch.cuda.synchronize()
        e_crm = float((crm.cpu() - ref_crm).abs().max() / ref_crm.abs().max())
        e_wav = float((wav.cpu() - ref_wav).abs().max())
        print(f"[smoke] precision={m._resolve_precision()} crm max-rel {e_crm:.2e} wav max-abs {e_wav:.2e}")
        assert e_crm < 1e-3 and e_wav < 1e-4, (e_crm, e_wav)
    print("[smoke] ok")


if __name__ == "__main__":
    build()
    if len(sys.argv) > 1 and sys.argv[1] == "smoke":
        smoke()
