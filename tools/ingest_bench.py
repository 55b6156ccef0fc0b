"""Developer timing probe for the GPU-side FASTA ingest: python tools/ingest_bench.py <config> [scale_n]
Writes the synthetic config as a 60-column multiFASTA (one record per fragment) to a tmpfs file, then times
  host   asgart_b200_prepare_files (C++ restatement of prepare_data on one host thread) + load_strand
  gpu    Context.ingest(path)  (file -> pinned staging -> HBM -> scans)  and  Context.ingest(bytes in memory)
and checks that both give the same strand, map and chunks."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asgart_b200 as ab  # noqa: E402


def fasta_bytes(g: np.ndarray, frags, width=60) -> np.ndarray:
    parts = []
    for name, pos, ln in frags:
        parts.append(np.frombuffer(f">{name} synthetic\n".encode(), dtype=np.uint8))
        seq = g[pos:pos + ln]
        full = ln // width * width
        body = np.empty((ln // width, width + 1), dtype=np.uint8)
        body[:, :width] = seq[:full].reshape(-1, width)
        body[:, width] = 10
        parts.append(body.reshape(-1))
        if ln > full:
            parts.append(seq[full:])
            parts.append(np.frombuffer(b"\n", dtype=np.uint8))
    return np.concatenate(parts)


def main():
    config = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    scale_n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    skip_masked = config == 2
    g, fr = ab.synth_genome(config, scale_n=scale_n, threads=os.cpu_count() or 8)
    fa = fasta_bytes(g, fr)
    path = f"/dev/shm/ingest_bench_c{config}.fa"
    fa.tofile(path)
    print(f"config C{config}: {len(g)} bp, {len(fr)} fragments, file {len(fa)} bytes", flush=True)
    res = {"config": config, "bp": int(len(g)), "file_bytes": int(len(fa))}
    t0 = time.perf_counter()
    hp = ab.prepare_data([path], skip_masked)
    res["host_prepare_s"] = round(time.perf_counter() - t0, 3)
    with ab.Context(0) as ctx:
        t0 = time.perf_counter()
        ctx.load_strand(hp.strand)
        res["host_load_strand_s"] = round(time.perf_counter() - t0, 3)
        want = np.array(hp.strand)
        for rep in range(3):
            ctx.reset_stats()
            t0 = time.perf_counter()
            gp = ctx.ingest([path], skip_masked)
            dt = time.perf_counter() - t0
            s = ctx.stats()
            res[f"gpu_file_rep{rep}"] = {"wall_s": round(dt, 3), "ms_h2d": round(s["ms_h2d"], 2), "ms_ingest": round(s["ms_ingest"], 2),
                                         "ms_pack": round(s["ms_pack"], 2), "ingest_GBps": round(s["ingest_bytes"] / max(s["ms_ingest"], 1e-9) / 1e6, 1)}
        assert gp.map == hp.map and gp.chunks == hp.chunks
        assert np.array_equal(ctx.download_strand(), want)
        for rep in range(2):
            ctx.reset_stats()
            t0 = time.perf_counter()
            gp = ctx.ingest([fa], skip_masked, names=[path])
            dt = time.perf_counter() - t0
            s = ctx.stats()
            res[f"gpu_mem_rep{rep}"] = {"wall_s": round(dt, 3), "ms_h2d": round(s["ms_h2d"], 2), "ms_ingest": round(s["ms_ingest"], 2)}
        assert gp.map == hp.map and gp.chunks == hp.chunks
    os.unlink(path)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
