"""Host->device copy ceiling of this box with N GPUs uploading at once (evidence tool; run under torchrun, one rank per GPU).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/h2d_probe.py

Each rank allocates 1.2 GB of pinned host memory (the bench's per-GPU input per step) after binding to its GPU's NUMA node,
then times, all ranks starting together: (a) ONE contiguous cudaMemcpyAsync of the buffer, (b) the 2-D copy the library's
host path issues (256 rows of 4.9 MB, cudaMemcpy2DAsync), (c) the same bytes as 7 contiguous pieces (the sub-chunk plan).
Also: the 2-D copy as two halves on two streams, and one contiguous copy per row (52.5 and 53.0 GB/s against 49.9 for the single 2-D
copy and 55.5 for one contiguous copy on this pool; splitting the library's uploads over two streams did not move its end-to-end
number, 24.2 G samples/s either way, so it stays one stream).
Rank 0 prints one JSON line with per-rank and aggregate GB/s: if (a) reaches what the bench's e2e leg reaches, the limit at
N = 8 is the host's memory system / PCIe topology, not the copy pattern."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dy4_b200 import shard


def main():
    rank, local, world = shard.init_process_group()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    rows, row_bytes = 256, 48 * 102400
    h = torch.empty((rows, row_bytes), dtype=torch.uint8).pin_memory()
    h.fill_(7)
    d = torch.empty((rows, row_bytes), dtype=torch.uint8, device=dev)
    d_wide = torch.empty((rows, row_bytes + 4096), dtype=torch.uint8, device=dev)      # a staging buffer with another row stride
    st = torch.cuda.Stream(dev)
    pieces = [1, 2, 4, 8, 16, 16, 1]                                                    # blocks per sub-chunk of a 48-block call

    st2 = torch.cuda.Stream(dev)

    def run(kind):
        if kind.startswith("rows_2d_x2"):                                               # the same 2-D copy as two halves on two streams (two copy engines)
            half = rows // 2
            with torch.cuda.stream(st):
                d_wide[:half, :row_bytes].copy_(h[:half], non_blocking=True)
            with torch.cuda.stream(st2):
                d_wide[half:, :row_bytes].copy_(h[half:], non_blocking=True)
            st.synchronize(); st2.synchronize()
            return
        if kind == "cols_2d_x2":                                                        # two streams, each half of every row's bytes
            hb = row_bytes // 2
            with torch.cuda.stream(st):
                d_wide[:, :hb].copy_(h[:, :hb], non_blocking=True)
            with torch.cuda.stream(st2):
                d_wide[:, hb:row_bytes].copy_(h[:, hb:], non_blocking=True)
            st.synchronize(); st2.synchronize()
            return
        with torch.cuda.stream(st):
            if kind == "contiguous":
                d.copy_(h, non_blocking=True)
            elif kind == "rows_2d":
                d_wide[:, :row_bytes].copy_(h, non_blocking=True)                       # strided destination: one 2-D copy
            elif kind == "rows_1d":                                                     # one contiguous copy per row (256 x 4.9 MB)
                for r in range(rows):
                    d_wide[r, :row_bytes].copy_(h[r], non_blocking=True)
            elif kind == "subchunk_rows_1d":                                            # one contiguous copy per row and sub-chunk (3 x 256 x 1.6 MB)
                for b0 in (0, 16, 32):
                    for r in range(rows):
                        d_wide[r, b0 * 102400:(b0 + 16) * 102400].copy_(h[r, b0 * 102400:(b0 + 16) * 102400], non_blocking=True)
            else:
                b0 = 0
                for nb in pieces:
                    d_wide[:, b0 * 102400:(b0 + nb) * 102400].copy_(h[:, b0 * 102400:(b0 + nb) * 102400], non_blocking=True)
                    b0 += nb
        st.synchronize()

    res = {}
    for kind in ("contiguous", "rows_2d", "rows_2d_x2", "cols_2d_x2", "rows_1d", "subchunk_rows_1d", "subchunks_2d"):
        run(kind)
        best = 1e9
        for _ in range(5):
            shard.barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            run(kind)
            dt = shard.max_over_ranks(time.perf_counter() - t0, dev)                   # the slowest rank ends the step
            best = min(best, dt)
        res[kind] = {"seconds": round(best, 5), "GBps_per_gpu": round(rows * row_bytes / best / 1e9, 2),
                     "GBps_aggregate": round(world * rows * row_bytes / best / 1e9, 2)}
    node = shard.bind_to_gpu_numa_node(local)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "bytes_per_gpu": rows * row_bytes, "host_cpus": os.cpu_count(), "numa_node_rank0": node, "copies": res}))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
