"""Staging gather under memory contention, CPU only (no GPU needed): P processes at once, each gathering its own 256 x 3 s
list (49 MB, sources rotating over 8 lists so that nothing stays in cache) through vp_host_stage_h2d in gather-only mode,
with memcpy and with streaming stores.  Dev tool for DESIGN 8.5; the absolute numbers are this container's, not a B200
host's."""
import ctypes as C
import multiprocessing as mp
import sys
import time

sys.path.insert(0, '.')
import numpy as np


def worker(rank, nproc, threads, streaming, barrier, out):
    from mvector import _lib as L
    lib = L.lib()
    lib.vp_host_gather_streaming(streaming)
    B, lmax, npool = 256, 48000, 8
    pools = [[np.full(lmax, i + j + rank, dtype=np.float32) for i in range(B)] for j in range(npool)]
    P = [np.fromiter((w.__array_interface__['data'][0] for w in p), dtype=np.uint64, count=B) for p in pools]
    lens = np.full(B, lmax, dtype=np.int32)
    dst = np.zeros((B, lmax), dtype=np.float32)

    def call(k):
        rc = lib.vp_host_stage_h2d(C.c_void_p(P[k % npool].ctypes.data), C.c_void_p(lens.ctypes.data), B, lmax,
                                   C.c_void_p(dst.ctypes.data), C.c_void_p(), 4, threads, C.c_void_p())
        assert rc == 0
    for k in range(3):
        call(k)
    barrier.wait()
    t0 = time.perf_counter()
    n = 16
    for k in range(n):
        call(k)
    out[rank] = (time.perf_counter() - t0) / n * 1e3
    assert dst[5, 7] == pools[(n - 1) % npool][5][7]


if __name__ == '__main__':
    import __graft_entry__ as ge
    ge.build()
    from mvector import _lib as L
    print('streaming stores available:', L.lib().vp_host_gather_streaming(1) == 1)
    for nproc, threads in ((1, 1), (1, 8), (2, 4), (4, 2), (8, 1)):
        for streaming in (0, 1):
            barrier = mp.Barrier(nproc)
            out = mp.Array('d', nproc)
            ps = [mp.Process(target=worker, args=(r, nproc, threads, streaming, barrier, out)) for r in range(nproc)]
            [p.start() for p in ps]
            [p.join() for p in ps]
            print(f'{nproc} process(es) x {threads} thread(s), {"streaming" if streaming else "memcpy   "}: '
                  f'ms per 49 MB per process {[round(x, 2) for x in out]}', flush=True)
