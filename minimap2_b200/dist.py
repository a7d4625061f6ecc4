"""Multi-GPU plumbing (one process per GPU, torch.distributed): reads shard across ranks with no collective on the
per-read path; the index is built/loaded on rank 0 and its device arrays are broadcast once (NCCL over NVLink/NVSwitch on
GPUs, gloo in the CPU tests) -- SURVEY 8(e). Results are merged in input order on rank 0."""
import ctypes as C
import numpy as np


def shard_bounds(qlens, world):
    """Contiguous shards balanced by bases (the same rule the in-process group scheduler uses, map.cu mm_map_batch)."""
    qlens = np.asarray(qlens, dtype=np.int64)
    total = int(np.maximum(qlens, 0).sum())
    cut = [0] * (world + 1)
    acc, g = 0, 1
    for i, l in enumerate(qlens):
        if g >= world:
            break
        acc += max(int(l), 0)
        while g < world and acc >= total * g // world:
            cut[g] = i + 1
            g += 1
    for k in range(g, world):
        cut[k] = len(qlens)
    cut[world] = len(qlens)
    return cut


def gather_in_order(local_items, cut, rank, world, dist, dst=0):
    """Ordered merge of per-rank result lists (the ordering contract of kthread.c:107-112): rank r owns reads cut[r]:cut[r+1]."""
    out = [None] * world if rank == dst else None
    dist.gather_object(local_items, out, dst=dst)
    if rank != dst:
        return None
    merged = []
    for r in range(world):
        assert len(out[r]) == cut[r + 1] - cut[r]
        merged.extend(out[r])
    return merged


class IdxDesc(C.Structure):  # mmb_idx_desc_t (index.cu)
    _fields_ = [("ptr", C.c_void_p * 5), ("bytes", C.c_uint64 * 5), ("n_keys", C.c_int64), ("n_pos", C.c_int64),
                ("tab_bits", C.c_int32), ("w", C.c_int32), ("k", C.c_int32), ("b", C.c_int32), ("flag", C.c_int32),
                ("n_seq", C.c_uint32), ("sum_len", C.c_uint64)]


class _DevView:
    """exposes a raw device pointer to torch through __cuda_array_interface__"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


def broadcast_index(idx, rank, src=0):
    """Rank `src` holds a built index (ctypes pointer); every other rank passes None and receives a replica.
    Returns (idx_pointer, keepalive) -- keepalive holds the torch buffers backing an adopted index."""
    import torch
    import torch.distributed as dist
    from . import api
    L = api._setup()
    L.mmb_idx_export.argtypes = [C.POINTER(api.Idx), C.POINTER(IdxDesc)]
    L.mmb_idx_adopt.restype = C.POINTER(api.Idx)
    L.mmb_idx_adopt.argtypes = [C.POINTER(IdxDesc), C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p]
    L.mmb_idx_cnt_sorted.restype = C.c_void_p
    L.mmb_idx_cnt_sorted.argtypes = [C.POINTER(api.Idx), C.POINTER(C.c_uint64)]
    d = IdxDesc()
    meta = [None]
    if rank == src:
        L.mmb_idx_export(idx, C.byref(d))
        nb = C.c_uint64(0)
        cs = L.mmb_idx_cnt_sorted(idx, C.byref(nb))
        names = [idx.contents.seq[i].name.decode() for i in range(idx.contents.n_seq)]
        lens = [int(idx.contents.seq[i].len) for i in range(idx.contents.n_seq)]
        meta = [dict(bytes=list(d.bytes) + [int(nb.value)], n_keys=d.n_keys, n_pos=d.n_pos, tab_bits=d.tab_bits, w=d.w, k=d.k, b=d.b, flag=d.flag,
                     n_seq=d.n_seq, sum_len=d.sum_len, names=names, lens=lens)]
    dist.broadcast_object_list(meta, src=src)
    m = meta[0]
    bufs = []
    for i in range(6):
        if rank == src:
            ptr = d.ptr[i] if i < 5 else cs
            t = torch.as_tensor(_DevView(ptr, m["bytes"][i]), device="cuda")
        else:
            t = torch.empty(m["bytes"][i], dtype=torch.uint8, device="cuda")
        dist.broadcast(t, src=src)
        bufs.append(t)
    if rank == src:
        return idx, bufs
    for i in range(5):
        d.ptr[i] = bufs[i].data_ptr(); d.bytes[i] = m["bytes"][i]
    d.n_keys, d.n_pos, d.tab_bits, d.w, d.k, d.b, d.flag, d.n_seq, d.sum_len = (m["n_keys"], m["n_pos"], m["tab_bits"], m["w"], m["k"], m["b"],
                                                                               m["flag"], m["n_seq"], m["sum_len"])
    names = (C.c_char_p * m["n_seq"])(*[s.encode() for s in m["names"]])
    lens = np.asarray(m["lens"], dtype=np.uint32)
    new_idx = L.mmb_idx_adopt(C.byref(d), names, lens.ctypes.data, bufs[5].data_ptr())
    return new_idx, bufs
