// fasta_ingest.cuh — FASTA / multiFASTA bytes -> normalised strand, record table and long N-runs, on the device.
//
// Replaces, for a file already in memory, the host-side read_fasta + find_chunks_to_process of prepare_data
// (src/bin/asgart.rs:278-366). The record semantics are those of the bio FASTA reader the reference
// calls (src/bin/asgart.rs:282-290): a line whose first byte is '>' opens a record, every other line is a sequence line
// and contributes its bytes with the trailing white space cut off (`trim_end`); white space INSIDE a sequence line stays
// (and becomes 'N' below, exactly as the reference's normalisation treats any byte outside ATGCN).
//
// Three streaming passes over the file bytes, 32 bytes (two 128-bit loads) per thread, 8 KB per block, so the work does
// not depend on line lengths (a one-line 250 Mbp record costs what 4 M lines of 60 columns cost). Two one-bit state
// machines run over the bytes:
//   header  (forward)   '>' as first byte of a line sets it, '\n' clears it: a byte lies on a header line iff it is set
//   rest    (backward)  '\n' (or the end of the file) sets it, a non-white byte clears it: a blank byte is trailing white
//                       space iff it is set
// For both, a stretch of bytes is summarised by its last (first) event, and "latest event wins" is associative:
//   pass 1  per 8 KB tile: last header event, first rest event                     -> tiny tile-level scans (scan.cuh)
//   pass 2  per tile, with the states entering it: kept bytes, header starts, and the state of the N-run counter
//           (kept bytes since the last kept non-N byte)                             -> one tile-level scan of the triple
//   pass 3  compaction + normalisation of the kept bytes (src/bin/asgart.rs:291-301) staged in shared memory and written
//           with 128-bit stores; one (file offset, strand position) pair per record; an N-run longer than 5000
//           (src/bin/asgart.rs:326,336) is reported by the first non-N base after it (there are at most n/5001 of them)
// Algorithmic bytes: 3 reads per file byte + 1 write per kept byte.
#pragma once
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "fasta_core.h"
#include "scan.cuh"

namespace ab200 {

struct IngestPiece {
    DevBuf<u8> strand;               // kept, normalised bytes of one file (no '$')
    u64 kept = 0;
    std::vector<u64> rec_off;        // file offset of each record's '>'
    std::vector<u64> rec_pos;        // strand position (file-local) at which the record's sequence starts
    std::vector<u64> run_start, run_len;   // maximal N-runs longer than kLongNRun, ascending, file-local strand coordinates
    std::vector<std::string> names;  // record ids (header up to the first white space), filled by the caller
};

// Tile-level prefixes, all with commutative operators (the generic scans of scan.cuh reduce in strided order):
//   sums  kept bytes, header starts
//   max   strand position just after the last kept non-N byte (0: none so far). The N-run counter at any point is
//         (kept bytes so far) - (that position).
struct FiSum {
    u64 kept, headers;
    FiSum() = default;
    __host__ __device__ explicit FiSum(int) : kept(0), headers(0) {}
};
struct FiSumOp {
    __device__ __forceinline__ FiSum operator()(const FiSum& x, const FiSum& y) const {
        FiSum r(0);
        r.kept = x.kept + y.kept;
        r.headers = x.headers + y.headers;
        return r;
    }
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned lanemask_gt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_gt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ void fi_load(const u8* __restrict__ d_file, u64 n, FiThread& t) {
    t.base = u64(blockIdx.x) * kFiTile + u64(threadIdx.x) * kFiBytes;
    t.valid = t.base >= n ? 0 : int(n - t.base < u64(kFiBytes) ? n - t.base : u64(kFiBytes));
    t.vmask = t.valid == kFiBytes ? 0xffffffffu : ((1u << t.valid) - 1u);
    if (t.valid == kFiBytes) {
        const uint4* p = reinterpret_cast<const uint4*>(d_file + t.base);
        const uint4 v0 = __ldg(p), v1 = __ldg(p + 1);
        t.w[0] = v0.x; t.w[1] = v0.y; t.w[2] = v0.z; t.w[3] = v0.w;
        t.w[4] = v1.x; t.w[5] = v1.y; t.w[6] = v1.z; t.w[7] = v1.w;
    } else {
#pragma unroll
        for (int q = 0; q < kFiBytes / 4; ++q) {
            u32 x = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r) { const int j = 4 * q + r; x |= u32(j < t.valid ? d_file[t.base + j] : u8('\n')) << (8 * r); }
            t.w[q] = x;
        }
    }
    fi_classify(t, (t.base == 0 || t.valid == 0) ? 1u : u32(d_file[t.base - 1] == '\n'));
}

// states entering this thread from the left (header) and from the right (rest), given those entering the tile;
// also the tile's own summary (last header event, first rest event) in tile_h / tile_r
__device__ __forceinline__ void fi_block_states(const FiThread& t, u32 tile_h_in, u32 tile_r_in, u32& h_in, u32& r_in,
                                                u32& tile_h, u32& tile_r, u32* sm /* 2 * kFiWarps */) {
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const unsigned hb = __ballot_sync(0xffffffffu, t.h_last != FI_NONE), hs = __ballot_sync(0xffffffffu, t.h_last == FI_SET);
    const unsigned rb = __ballot_sync(0xffffffffu, t.r_first != FI_NONE), rs = __ballot_sync(0xffffffffu, t.r_first == FI_SET);
    if (lane == 0) {
        sm[warp] = hb ? (((hs >> (31 - __clz(hb))) & 1u) ? FI_SET : FI_CLR) : FI_NONE;
        sm[kFiWarps + warp] = rb ? (((rs >> (__ffs(rb) - 1)) & 1u) ? FI_SET : FI_CLR) : FI_NONE;
    }
    __syncthreads();
    u32 wh = tile_h_in, wr = tile_r_in;
    for (unsigned w = 0; w < warp; ++w) if (sm[w] != FI_NONE) wh = sm[w] == FI_SET;
    for (unsigned w = kFiWarps - 1; w > warp; --w) if (sm[kFiWarps + w] != FI_NONE) wr = sm[kFiWarps + w] == FI_SET;
    const unsigned ml = hb & lanemask_lt(), mg = rb & lanemask_gt();
    h_in = ml ? ((hs >> (31 - __clz(ml))) & 1u) : wh;
    r_in = mg ? ((rs >> (__ffs(mg) - 1)) & 1u) : wr;
    tile_h = FI_NONE; tile_r = FI_NONE;
    for (int w = 0; w < kFiWarps; ++w) if (sm[w] != FI_NONE) tile_h = sm[w];
    for (int w = kFiWarps - 1; w >= 0; --w) if (sm[kFiWarps + w] != FI_NONE) tile_r = sm[kFiWarps + w];
    __syncthreads();
}

__global__ void __launch_bounds__(kFiThreads) fi_events_kernel(const u8* __restrict__ d_file, u64 n, u8* __restrict__ tile_ev) {
    __shared__ u32 sm[2 * kFiWarps];
    FiThread t;
    fi_load(d_file, n, t);
    u32 h_in, r_in, th, tr;
    fi_block_states(t, 0, 1, h_in, r_in, th, tr, sm);
    if (threadIdx.x == 0) tile_ev[blockIdx.x] = u8(th | (tr << 2));
}

__global__ void __launch_bounds__(kFiThreads) fi_count_kernel(const u8* __restrict__ d_file, u64 n, const u8* __restrict__ tile_in,
                                                             bool skip_masked, FiSum* __restrict__ tile_sum, u32* __restrict__ tile_base) {
    __shared__ u32 sm[2 * kFiWarps];
    __shared__ u64 s64[32];
    FiThread t;
    fi_load(d_file, n, t);
    const u32 tin = tile_in[blockIdx.x];
    u32 h_in, r_in, th, tr, upto;
    fi_block_states(t, tin & 1u, (tin >> 1) & 1u, h_in, r_in, th, tr, sm);
    const u32 keep = fi_keep_mask(t, h_in, r_in);
    u64 packed, total, last;
    fi_thread_counts(t, keep, fi_base_mask(t, skip_masked, nullptr), packed, upto);
    const u64 exc = block_exclusive_scan(packed, SumOp(), total, s64);
    block_exclusive_scan(upto ? u64((exc & 0xffffffffull) + upto) : u64(0), MaxOp(), last, s64);
    if (threadIdx.x == 0) {
        FiSum ts(0);
        ts.kept = total & 0xffffffffull;
        ts.headers = total >> 32;
        tile_sum[blockIdx.x] = ts;
        tile_base[blockIdx.x] = u32(last);      // kept bytes of the tile up to and including its last base (0: none)
    }
}

__global__ void __launch_bounds__(kFiThreads) fi_emit_kernel(const u8* __restrict__ d_file, u64 n, const u8* __restrict__ tile_in,
                                                            const FiSum* __restrict__ tile_pre, const u64* __restrict__ tile_last,
                                                            bool skip_masked, u8* __restrict__ strand, u64* __restrict__ rec,
                                                            u64* __restrict__ run_cnt, u64* __restrict__ runs, u64 run_cap) {
    __shared__ u32 sm[2 * kFiWarps];
    __shared__ u64 s64[32];
    __shared__ __align__(16) u8 stage[kFiTile + 16];
    FiThread t;
    fi_load(d_file, n, t);
    const u32 tin = tile_in[blockIdx.x];
    u32 h_in, r_in, th, tr, upto;
    fi_block_states(t, tin & 1u, (tin >> 1) & 1u, h_in, r_in, th, tr, sm);
    const u32 keep = fi_keep_mask(t, h_in, r_in);
    u32 norm[kFiBytes / 4];
    const u32 base = fi_base_mask(t, skip_masked, norm);
    u64 packed, total, last_total;
    fi_thread_counts(t, keep, base, packed, upto);
    const u64 exc = block_exclusive_scan(packed, SumOp(), total, s64);
    const u64 last_in_tile = block_exclusive_scan(upto ? u64((exc & 0xffffffffull) + upto) : u64(0), MaxOp(), last_total, s64);
    const FiSum tile0 = tile_pre[blockIdx.x];
    const u64 out0 = tile0.kept;
    const u32 mis = u32(out0 & 15u);     // the strand buffer is 256-byte aligned: same misalignment in the staging area
    const u64 outpos = out0 + (exc & 0xffffffffull);
    if (t.heads) {                       // one (file offset, strand position) pair per record
        u64 hidx = tile0.headers + (exc >> 32);
        for (u32 m = t.heads; m; m &= m - 1) {
            const int j = __ffs(m) - 1;
            rec[2 * hidx] = t.base + j;
            rec[2 * hidx + 1] = outpos + __popc(keep & ((1u << j) - 1u));
            ++hidx;
        }
    }
    const u32 kb = keep & base;
    if (kb) {                            // an N-run longer than 5000 is reported by the first base after it
        const u64 after_last = last_in_tile ? out0 + last_in_tile : tile_last[blockIdx.x];
        const u64 first_base = outpos + __popc(keep & ((1u << (__ffs(kb) - 1)) - 1u));
        const u64 len = first_base - after_last;
        if (len > kLongNRun) {
            const u64 slot = atomicAdd(reinterpret_cast<unsigned long long*>(run_cnt), 1ull);
            if (slot < run_cap) { runs[2 * slot] = after_last; runs[2 * slot + 1] = len; }
        }
    }
    u32 at = mis + u32(outpos - out0);
    if (keep == 0xffffffffu && (at & 3u) == 0) {
#pragma unroll
        for (int q = 0; q < kFiBytes / 4; ++q) *reinterpret_cast<u32*>(stage + at + 4 * q) = norm[q];
    } else {
#pragma unroll
        for (int j = 0; j < kFiBytes; ++j)
            if ((keep >> j) & 1u) stage[at++] = u8(norm[j >> 2] >> ((j & 3) * 8));
    }
    __syncthreads();
    const u32 cnt = u32(total & 0xffffffffull);
    const u32 head = min(cnt, (16u - mis) & 15u);
    const u32 nvec = (cnt - head) / 16u;
    if (threadIdx.x < head) strand[out0 + threadIdx.x] = stage[mis + threadIdx.x];
    for (u32 v = threadIdx.x; v < nvec; v += kFiThreads)
        *reinterpret_cast<uint4*>(strand + out0 + head + 16u * v) = *reinterpret_cast<const uint4*>(stage + mis + head + 16u * v);
    for (u32 i = head + 16u * nvec + threadIdx.x; i < cnt; i += kFiThreads) strand[out0 + i] = stage[mis + i];
}
#endif  // __CUDACC__

inline void ingest_fasta_device(const u8* d_file, u64 n, bool skip_masked, IngestPiece& out, cudaStream_t stream) {
    out.kept = 0;
    out.rec_off.clear(); out.rec_pos.clear(); out.run_start.clear(); out.run_len.clear();
    if (n == 0) return;
    const u64 tiles = ceil_div(n, kFiTile);
    DevBuf<u8> tile_ev(tiles, stream), tile_in(tiles, stream);
    fi_events_kernel<<<unsigned(tiles), kFiThreads, 0, stream>>>(d_file, n, tile_ev.p);
    KERNEL_CHECK();
    count_launch();
    {   // states entering each tile: the latest header event before it (none: not on a header line), the first rest event
        // after it (none: the end of the file). "Latest event" = running maximum of (position, value) keys.
        const u8* ev = tile_ev.p;
        u8* tin = tile_in.p;
        auto h_key = [=] __device__(u64 i) -> u64 { const u32 e = ev[i] & 3u; return e ? (((i + 1) << 1) | u64(e == FI_SET)) : 0; };
        auto h_out = [=] __device__(u64 i, u64 exc, u64) { tin[i] = u8(exc ? (exc & 1) : 0); };
        device_scan<u64, MaxOp>(h_key, h_out, tiles, (u64*)nullptr, stream);
        auto r_key = [=] __device__(u64 i) -> u64 { const u32 e = (ev[tiles - 1 - i] >> 2) & 3u; return e ? (((i + 1) << 1) | u64(e == FI_SET)) : 0; };
        auto r_out = [=] __device__(u64 i, u64 exc, u64) { tin[tiles - 1 - i] |= u8((exc ? (exc & 1) : 1) << 1); };
        device_scan<u64, MaxOp>(r_key, r_out, tiles, (u64*)nullptr, stream);
    }
    DevBuf<FiSum> tile_sum(tiles, stream), tile_pre(tiles, stream), d_total(1, stream);
    DevBuf<u32> tile_base(tiles, stream);
    DevBuf<u64> tile_last(tiles, stream), d_last(1, stream);
    fi_count_kernel<<<unsigned(tiles), kFiThreads, 0, stream>>>(d_file, n, tile_in.p, skip_masked, tile_sum.p, tile_base.p);
    KERNEL_CHECK();
    count_launch();
    {
        const FiSum* sum = tile_sum.p;
        FiSum* pre = tile_pre.p;
        const u32* base = tile_base.p;
        u64* last = tile_last.p;
        auto in = [=] __device__(u64 i) -> FiSum { return sum[i]; };
        auto wr = [=] __device__(u64 i, const FiSum& exc, const FiSum&) { pre[i] = exc; };
        device_scan<FiSum, FiSumOp>(in, wr, tiles, d_total.p, stream);
        auto lin = [=] __device__(u64 i) -> u64 { return base[i] ? pre[i].kept + base[i] : 0; };
        auto lwr = [=] __device__(u64 i, u64 exc, u64) { last[i] = exc; };
        device_scan<u64, MaxOp>(lin, lwr, tiles, d_last.p, stream);
    }
    FiSum tot(0);
    u64 after_last = 0;
    CUDA_CHECK(cudaMemcpyAsync(&tot, d_total.p, sizeof tot, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaMemcpyAsync(&after_last, d_last.p, sizeof after_last, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    out.kept = tot.kept;
    const u64 records = tot.headers;
    const u64 cap = out.kept / (kLongNRun + 1) + 1;
    out.strand.alloc(out.kept, stream);
    DevBuf<u64> d_rec(2 * records, stream), d_runs(2 * cap + 1, stream);
    CUDA_CHECK(cudaMemsetAsync(d_runs.p, 0, sizeof(u64), stream));
    fi_emit_kernel<<<unsigned(tiles), kFiThreads, 0, stream>>>(d_file, n, tile_in.p, tile_pre.p, tile_last.p, skip_masked, out.strand.p,
                                                               d_rec.p, d_runs.p, d_runs.p + 1, cap);
    KERNEL_CHECK();
    count_launch();
    u64 h_cnt = 0;
    std::vector<u64> h(2 * records);
    if (records) CUDA_CHECK(cudaMemcpyAsync(h.data(), d_rec.p, h.size() * sizeof(u64), cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaMemcpyAsync(&h_cnt, d_runs.p, sizeof h_cnt, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    out.rec_off.resize(records); out.rec_pos.resize(records);
    for (u64 r = 0; r < records; ++r) { out.rec_off[r] = h[2 * r]; out.rec_pos[r] = h[2 * r + 1]; }
    h_cnt = std::min(h_cnt, cap);
    std::vector<u64> hr(2 * h_cnt);
    if (h_cnt) {
        CUDA_CHECK(cudaMemcpyAsync(hr.data(), d_runs.p + 1, hr.size() * sizeof(u64), cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    std::vector<std::pair<u64, u64>> v(h_cnt);
    for (u64 i = 0; i < h_cnt; ++i) v[i] = {hr[2 * i], hr[2 * i + 1]};
    if (out.kept - after_last > kLongNRun) v.push_back({after_last, out.kept - after_last});   // the run that ends the strand
    std::sort(v.begin(), v.end());
    for (auto& p : v) { out.run_start.push_back(p.first); out.run_len.push_back(p.second); }
}

// chunks_to_process and fragment table of one ingested file; `off` = strand position of the file's first base
inline void ingest_chunks(const IngestPiece& p, u64 off, std::vector<asgart_b200_chunk>& chunks,
                          std::vector<u64>& frag_pos, std::vector<u64>& frag_len) {
    chunks_from_runs(p.rec_pos, p.kept, p.run_start, p.run_len, off, chunks, frag_pos, frag_len);
}

}  // namespace ab200
