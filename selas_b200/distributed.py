"""Multi-GPU plumbing for the one place the path needs it: the scene is replicated on every GPU,
sample passes are sharded across ranks, and ONE reduce of the float accumulation buffers merges the films.

torch.distributed is used as plumbing only (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations


def shard_samples(total_spp: int, rank: int, world: int):
    """Contiguous split of sample indices [0, total_spp) over `world` ranks; returns (first_sample, count).
    The first total_spp % world ranks get one extra sample, so every sample index is rendered exactly once."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(total_spp), world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def shard_iterations(total_iterations: int, rank: int, world: int):
    """VCM shards whole iterations (each pass has its own light paths, hash grid and radius; SelasVcmSettings.firstIteration):
    returns (first_iteration, count) with the same contiguous split as shard_samples."""
    return shard_samples(total_iterations, rank, world)


def reduce_film(film, dst: int = 0):
    """Sum the per-rank accumulation buffers onto rank `dst` (the path's only collective); no-op without a process group."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film


class FilmReducer:
    """The path's one collective through the product's C ABI: selas_cuda_reduce_film = ncclReduce(SUM, float32) onto `root` over a
    raw ncclComm_t owned by this object.  The communicator's unique id is created on rank 0 and handed to the other ranks through
    the already-initialised torch.distributed group (plumbing only); NCCL itself is the library the process already has loaded."""

    def __init__(self, ctx, rank: int, world: int):
        import ctypes as C

        import torch
        import torch.distributed as dist

        class UniqueId(C.Structure):
            _fields_ = [("internal", C.c_byte * 128)]  # NCCL_UNIQUE_ID_BYTES, nccl.h

        self.ctx, self.rank, self.world = ctx, rank, world
        self.nccl = C.CDLL("libnccl.so.2")
        uid = UniqueId()
        if rank == 0:
            rc = self.nccl.ncclGetUniqueId(C.byref(uid))
            if rc:
                raise RuntimeError(f"ncclGetUniqueId failed: {rc}")
        dev = torch.device("cuda", torch.cuda.current_device())
        t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).to(dev)
        dist.broadcast(t, src=0)
        C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
        self.comm = C.c_void_p()
        self.nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
        rc = self.nccl.ncclCommInitRank(C.byref(self.comm), world, uid, rank)
        if rc:
            raise RuntimeError(f"ncclCommInitRank failed: {rc}")

    def reduce(self, film, dst: int = 0, stream: int = 0):
        """In-place sum of the float32 device tensor `film` onto rank dst, on `stream` (0 = the context's stream)."""
        import ctypes as C

        self.ctx.check(self.ctx.lib.selas_cuda_reduce_film(self.ctx.handle, self.comm, C.c_void_p(film.data_ptr()), C.c_size_t(film.numel()),
                                                           C.c_int(dst), C.c_void_p(stream)))
        return film

    def close(self):
        if getattr(self, "comm", None) is not None and self.comm.value:
            self.nccl.ncclCommDestroy.argtypes = [__import__("ctypes").c_void_p]
            self.nccl.ncclCommDestroy(self.comm)
            self.comm = None
