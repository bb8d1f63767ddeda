"""selas_cuda_reduce_film: the path's only collective through the C ABI, with NCCL bound at run time.
One process can only hold a 1-rank communicator here; the N-rank form runs in bench.py / scripts/bench_vcm.py under torchrun
(profiles/: 2- and 8-GPU results)."""
import ctypes as C

import numpy as np
import pytest

from selas_b200 import capi

pytestmark = pytest.mark.gpu


class UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


def test_reduce_film_single_rank(ctx):
    import torch

    nccl = C.CDLL("libnccl.so.2")
    uid = UniqueId()
    assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    assert nccl.ncclCommInitRank(C.byref(comm), 1, uid, 0) == 0
    film = torch.arange(3 * 64 * 48, dtype=torch.float32, device="cuda") * 0.25
    want = film.clone()
    stream = torch.cuda.current_stream().cuda_stream
    rc = ctx.lib.selas_cuda_reduce_film(ctx.handle, comm, C.c_void_p(film.data_ptr()), C.c_size_t(film.numel()), C.c_int(0), C.c_void_p(stream))
    assert rc == 0, ctx.lib.selas_cuda_last_error(ctx.handle)
    torch.cuda.synchronize()
    assert torch.equal(film, want)  # one rank: the sum is the buffer itself
    # argument checking
    assert ctx.lib.selas_cuda_reduce_film(ctx.handle, None, C.c_void_p(film.data_ptr()), C.c_size_t(film.numel()), C.c_int(0), None) == 2
    assert b"communicator" in ctx.lib.selas_cuda_last_error(ctx.handle)
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    nccl.ncclCommDestroy(comm)
