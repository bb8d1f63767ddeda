#!/usr/bin/env python
"""Where the Ptex lookup's time goes: the same number of lookups restricted to kernel placements that stay inside the face,
cross one edge, or cross a corner (8x8 uint8 faces, random face ids; --jitter 1: resolutions mixed 8 / 16 per axis).  Prints one JSON
object.  python scripts/ptex_classes.py [--grid 800] [--lookups 16000000] [--jitter 1]"""
import argparse
import json
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from selas_b200 import capi  # noqa: E402
from tests import oracle_binding  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=800)
    ap.add_argument("--lookups", type=int, default=16_000_000)
    ap.add_argument("--ptx", default="", help="reuse / keep the authored .ptx at this path (A/B runs)")
    ap.add_argument("--log2", type=int, default=3, help="base face resolution (log2, both axes)")
    ap.add_argument("--jitter", type=int, default=0, help="a third / a quarter of the faces get 2^jitter times the resolution along u / v "
                    "(oracle/ptex_harness.cpp), so that most kernels that leave their face meet another resolution")
    args = ap.parse_args()
    import torch

    ptexref = oracle_binding.load_ptexref()
    tmp = Path(args.ptx) if args.ptx else Path(tempfile.mkdtemp()) / "c.ptx"
    if not tmp.exists():
        ptexref.write_grid(tmp, args.grid, args.grid, args.log2, args.log2, args.jitter, 0, 3, 4, True, 0)
    ctx = capi.Context(0)
    tex = capi.Ptex(ctx, tmp)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    rng = np.random.default_rng(7)
    n = args.lookups
    nfaces = args.grid * args.grid
    classes = {"interior": ((0.3, 0.7), (0.3, 0.7)), "one_edge": ((0.0, 0.1), (0.3, 0.7)), "corner": ((0.0, 0.1), (0.9, 1.0)),
               "uniform": ((0.0, 1.0), (0.0, 1.0))}
    out = {}
    for order in ("random", "coherent"):
        face = rng.integers(0, nfaces, size=n).astype(np.int32) if order == "random" else (np.arange(n, dtype=np.int64) * nfaces // n).astype(np.int32)
        d_f = torch.from_numpy(face).to(dev)
        for name, (ur, vr) in classes.items():
            u = rng.uniform(*ur, size=n).astype(np.float32)
            v = rng.uniform(*vr, size=n).astype(np.float32)
            d_u, d_v = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev)
            d_o = torch.empty(3 * n, dtype=torch.float32, device=dev)
            ms = []
            with torch.cuda.stream(stream):
                for it in range(6):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    tex.eval_device(n, d_f.data_ptr(), d_u.data_ptr(), d_v.data_ptr(), d_o.data_ptr(), stream.cuda_stream)
                    e1.record(stream)
                    e1.synchronize()
                    ms.append(e0.elapsed_time(e1))
            m = float(np.median(ms[1:]))
            k = 200_000
            want = ptexref.eval(tmp, face[:k], u[:k], v[:k])
            got = d_o[: 3 * k].cpu().numpy().reshape(-1, 3)
            out[f"{order}_{name}"] = {"ms": m, "mlookups_per_s": n / m / 1e3, "max_abs_diff_vs_reference": float(np.abs(got - want).max())}
    print(json.dumps(out, indent=1))
    tex.close()


if __name__ == "__main__":
    main()
