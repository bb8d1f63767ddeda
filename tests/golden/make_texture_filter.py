#!/usr/bin/env python
"""Golden vectors for the texture filters (TextureLib/TextureFiltering.h:54-234), produced by the reference's own templates compiled in
place (oracle/_ref/libselas_ref.so: ref_texture_filter).  Run in the container that has /root/reference:
    python tests/golden/make_texture_filter.py        -> tests/golden/texture_filter.npz"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from selas_b200 import capi  # noqa: E402
from tests import inputs, oracle_binding  # noqa: E402

CASES = ((capi.FILTER_POINT, 0), (capi.FILTER_TRIANGLE, 0), (capi.FILTER_TRIANGLE, 2), (capi.FILTER_TRILINEAR, 0), (capi.FILTER_EWA, 0),
         (capi.FILTER_EWA_LEVEL, 1))


def main():
    ref = oracle_binding.load_ref()
    assert ref is not None, "oracle/_ref/libselas_ref.so missing: make -C oracle"
    out = {}
    for comps in (1, 3, 4):
        levels = inputs.mip_chain(32, comps, seed=30 + comps)
        st, d0, d1 = inputs.filter_queries(600, seed=40 + comps)
        for flt, level in CASES:
            out[f"c{comps}_f{flt}_l{level}"] = ref.texture_filter(levels, flt, st, d0, d1, level)
    np.savez_compressed(Path(__file__).resolve().parent / "texture_filter.npz", **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
