"""Runs the UNMODIFIED reference model (vendored by oracle/make_ref.py into oracle/_ref/) on the CPU.

TEST / BASELINE INFRASTRUCTURE ONLY: bench.py's `--impl reference` arm and `cpu_baseline` leg use it when oracle/_ref
exists (kind "reference"); otherwise they fall back to the oracle port (oracle/aarmvs_oracle.py, kind "port").
Recipe = SURVEY.md §8c: sys.path.append (never insert(0)), `.cuda()` patched to identity because models/module.py:94-96 and
models/drmvsnet.py:302-304 hard-code it, evidential head stubbed (no weights in the shipped checkpoints, raises for
D != 32), strict=False load.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF_DIR, "models", "drmvsnet.py"))


class _NoEvidential(torch.nn.Module):
    def forward(self, prob_volume, depth_values):
        return torch.zeros(1), None


def load_reference(state_dict, h, w, return_depth=True):
    """EMVSNet of the vendored reference with the shipped weights, on the CPU, eval mode."""
    if not available():
        raise RuntimeError("oracle/_ref is absent: run `python oracle/make_ref.py` where /root/reference exists")
    if REF_DIR not in sys.path:
        sys.path.append(REF_DIR)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    from models.drmvsnet import EMVSNet
    m = EMVSNet(disparity_level=32, image_scale=1.0, max_h=h, max_w=w, return_depth=return_depth)
    res = m.load_state_dict(state_dict, strict=False)
    assert not res.unexpected_keys and all(k.startswith("evidential.") for k in res.missing_keys)
    m.evidential = _NoEvidential()
    return m.eval()
