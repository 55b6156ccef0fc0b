// levenshtein.cuh — the ComputeScore step (src/bin/asgart.rs:98-111) on the GPU: identity of every duplicon from the
// unit-cost edit distance of its two arms, as ProtoSD::levenshtein computes it (src/structs.rs:439-452):
//   left_arm  = strand[left ..= left + left_length]            (inclusive range: left_length + 1 bytes)
//   right_arm = strand[right ..= right + right_length], reversed if `reversed`, THEN complemented if `complemented`
//   identity  = (100.0 * (1.0 - dist / max(left_length, right_length) as f64)) as f32        (f64 arithmetic, one cast)
// `bio::alignment::distance::levenshtein` (bio "*", Cargo.toml:13; not on disk) is the plain global edit distance, so any
// exact algorithm reproduces it. Here: Myers' bit-parallel recurrence in Hyyrö's block form for the global distance
// (vertical deltas initialised to +1, horizontal input +1 along the top row, score read at the bottom-right cell).
//
// Layout. The left arm is the "pattern" (rows), the right arm the "text" (columns). Rows are cut into strips of
// 32 lanes x 64 rows = 2048 rows; one warp sweeps one strip along an anti-diagonal: at step s lane w advances its 64-row
// block through column s - w, and hands (symbol of that column, horizontal delta at the block's bottom row) to lane w + 1
// with one shuffle. The horizontal deltas leaving the strip's last row go to a byte row in global memory, the next strip
// reads them as its top-row input. One block of kLevWarps warps per duplicon: warp k owns strips k, k + kLevWarps, ... and
// follows the warp of the strip above at a distance of one 32-column chunk (progress counters in shared memory), so a
// long pair is swept by up to kLevWarps strips at once. Jobs are sorted by work on the host, heaviest first.
// Nothing here is HBM-bound: per 64x1 cell block the inner loop is ~45 integer instructions.
#pragma once
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace ab200 {

constexpr int kLevWarps = 8;
constexpr u32 kLevStripRows = 32 * 64;

struct LevJob {
    u64 left, right;   // first byte of each arm in the strand
    u64 m, n;          // arm lengths in bytes (left_length + 1, right_length + 1)
    u64 bnd_off;       // this job's boundary rows inside the scratch buffer (rows x row_stride bytes)
    u64 row_stride;
    u32 sd_index;      // where the distance goes
    u32 flags;         // bit 0 reversed, bit 1 complemented
};

// strand byte -> symbol code 0..5 (A, C, G, T, N, '$'); anything else cannot occur in a normalised strand
__device__ __forceinline__ u32 lev_code(u8 c) {
    return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : c == '$' ? 5u : 4u;
}

// One column through one 64-row block (Hyyrö 2003; hc = horizontal delta + 1 at the block's top row, returns the same
// for the row `high` of the block).
__device__ __forceinline__ u32 lev_advance(u64& Pv, u64& Mv, u64 Eq, u32 hc, u64 high) {
    const u64 hneg = hc == 0 ? 1ull : 0ull, hpos = hc == 2 ? 1ull : 0ull;
    const u64 Xv = Eq | Mv;
    Eq |= hneg;
    const u64 Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
    u64 Ph = Mv | ~(Xh | Pv);
    u64 Mh = Pv & Xh;
    const u32 out = 1u + ((Ph & high) ? 1u : 0u) - ((Mh & high) ? 1u : 0u);
    Ph = (Ph << 1) | hpos;
    Mh = (Mh << 1) | hneg;
    Pv = Mh | ~(Xv | Ph);
    Mv = Ph & Xv;
    return out;
}

__global__ void __launch_bounds__(kLevWarps * 32) levenshtein_kernel(const u8* __restrict__ text, const LevJob* __restrict__ jobs,
                                                                     u8* __restrict__ bnd, u32* __restrict__ dist_out,
                                                                     u32* __restrict__ err) {
    __shared__ unsigned long long progress[kLevWarps];   // columns of boundary row k completed, counted over all rounds
    const LevJob job = jobs[blockIdx.x];
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (threadIdx.x < kLevWarps) progress[threadIdx.x] = 0;
    __syncthreads();
    const u64 m = job.m, n = job.n;
    const u64 strips = (m + kLevStripRows - 1) / kLevStripRows;
    const bool rev = job.flags & 1u, comp = job.flags & 2u;
    volatile unsigned long long* prog = progress;

    for (u64 t = warp; t < strips; t += kLevWarps) {
        const u64 row0 = t * kLevStripRows;
        const u32 rows = u32(min(u64(kLevStripRows), m - row0));
        const u32 nw = (rows + 63) / 64;                 // lanes with rows in this strip
        const bool last_strip = t + 1 == strips;
        // ---- Eq masks of this lane's 64 rows, built by ballots over coalesced loads of the pattern
        u64 eq0 = 0, eq1 = 0, eq2 = 0, eq3 = 0, eq4 = 0, eq5 = 0;
        for (u32 r = 0; r < 2 * nw; ++r) {
            const u64 row = row0 + u64(r) * 32 + lane;
            const u32 c = row < m ? lev_code(text[job.left + row]) : 7u;
            const u64 b0 = __ballot_sync(0xffffffffu, c == 0), b1 = __ballot_sync(0xffffffffu, c == 1),
                      b2 = __ballot_sync(0xffffffffu, c == 2), b3 = __ballot_sync(0xffffffffu, c == 3),
                      b4 = __ballot_sync(0xffffffffu, c == 4), b5 = __ballot_sync(0xffffffffu, c == 5);
            if (lane == (r >> 1)) {
                const int sh = (r & 1) * 32;
                eq0 |= b0 << sh; eq1 |= b1 << sh; eq2 |= b2 << sh; eq3 |= b3 << sh; eq4 |= b4 << sh; eq5 |= b5 << sh;
            }
        }
        // row of this block whose horizontal delta leaves it: bit 63, or the last row of the arm in its last block
        const u64 high = (lane + 1 == nw) ? (1ull << ((rows - 1) & 63u)) : (1ull << 63);
        u64 Pv = ~0ull, Mv = 0;
        u32 carry = 1;        // (symbol << 2) | (hout + 1) of this lane's latest column
        i64 score = 0;        // last strip, lane nw - 1: sum of the horizontal deltas along the arm's last row
        const u8* in_row = t > 0 ? bnd + job.bnd_off + ((t - 1) % kLevWarps) * job.row_stride : nullptr;
        u8* out_row = bnd + job.bnd_off + (t % kLevWarps) * job.row_stride;
        const u64 in_base = t > 0 ? ((t - 1) / kLevWarps) * n : 0;   // progress count of the producer when its round began
        const u64 out_base = (t / kLevWarps) * n;
        const u32 pw = u32((t + kLevWarps - 1) % kLevWarps);         // warp of the strip above

        // this lane's column of the 32-column chunk at cb: byte of the right arm and top-row delta, still undecoded so
        // that nothing waits on the loads before the chunk in hand has been computed
        auto load_feed = [&](u64 cb, u32& raw, u32& hc) {
            const u64 col = cb + lane;
            raw = 'N'; hc = 2u;                   // top row of the matrix: D[0][c] - D[0][c-1] = +1
            if (col >= n) return;                 // beyond the arm: never consumed by an active block
            raw = text[job.right + (rev ? n - 1 - col : col)];
            if (t > 0) {
                const u64 need = in_base + min(n, cb + 32);
                while (prog[pw] < need) __nanosleep(32);
                __threadfence_block();
                hc = __ldcg(in_row + col);
            }
        };
        const u64 col_steps = n + nw - 1;
        u32 raw_next, hc_next;
        load_feed(0, raw_next, hc_next);
        for (u64 cb = 0; cb < col_steps; cb += 32) {
            u32 c = lev_code(u8(raw_next));
            if (comp) {
                if (c == 5u) atomicOr(err, 1u);   // the reference's complement() panics on '$' (src/structs.rs:28-34)
                else if (c < 4u) c ^= 3u;         // A<->T, C<->G, N stays
            }
            const u32 feed = c << 2 | hc_next;
            // the next chunk's loads are in flight while this one is computed (they trail the strip above by one more chunk)
            if (cb + 32 < col_steps) load_feed(cb + 32, raw_next, hc_next);
            __syncwarp();
#pragma unroll
            for (u32 j = 0; j < 32; ++j) {
                const u32 above = __shfl_up_sync(0xffffffffu, carry, 1);
                const u32 top = __shfl_sync(0xffffffffu, feed, j);
                const u32 in = lane == 0 ? top : above;
                const u64 c = cb + j - lane;            // wraps for lanes that have not started: fails the test below
                if (lane < nw && c < n) {
                    const u32 sym = in >> 2;
                    const u64 Eq = sym == 0 ? eq0 : sym == 1 ? eq1 : sym == 2 ? eq2 : sym == 3 ? eq3 : sym == 4 ? eq4 : eq5;
                    const u32 hout = lev_advance(Pv, Mv, Eq, in & 3u, high);
                    carry = (sym << 2) | hout;
                    if (lane + 1 == nw) {
                        if (last_strip) score += i64(hout) - 1;
                        else __stcg(out_row + c, u8(hout));
                    }
                }
            }
            if (!last_strip) {   // columns < cb + 32 - (nw - 1) of the outgoing row are final
                __threadfence_block();
                __syncwarp();
                if (lane == 0) {
                    const u64 done = cb + 32 >= u64(nw - 1) ? min(n, cb + 32 - (nw - 1)) : 0;
                    prog[warp] = out_base + done;
                }
            }
        }
        if (last_strip && lane + 1 == nw) dist_out[job.sd_index] = u32(i64(m) + score);
    }
}

// identity = 100 * (1 - dist / max(ll, rl)) in f64, cast to f32 (src/structs.rs:451, src/bin/asgart.rs:108)
__global__ void lev_identity_kernel(asgart_b200_protosd* __restrict__ sds, const u32* __restrict__ dist, u64 n_sds) {
    const u64 j = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n_sds) return;
    const u64 mx = max(sds[j].left_length, sds[j].right_length);
    sds[j].identity = float(100.0 * (1.0 - double(dist[j]) / double(mx)));
}

struct LevStats {
    double ms = 0;
    u64 cells = 0, pairs = 0;
};

// d_sds[0..n_sds): identity filled in place. n1 = strand bytes incl. '$'. Returns 0, or a reason the reference itself would
// have panicked on: 1 = complement of '$' (an arm that ends on the terminator in a -C/-RC run), 2 = arm past the strand end.
inline int compute_score(const u8* d_text, u64 n1, asgart_b200_protosd* d_sds, u64 n_sds, cudaStream_t s, LevStats* stats) {
    if (n_sds == 0) return 0;
    std::vector<asgart_b200_protosd> h(n_sds);
    CUDA_CHECK(cudaMemcpyAsync(h.data(), d_sds, n_sds * sizeof(asgart_b200_protosd), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    std::vector<LevJob> jobs(n_sds);
    for (u64 j = 0; j < n_sds; ++j) {
        const asgart_b200_protosd& sd = h[j];
        if (sd.left + sd.left_length >= n1 || sd.right + sd.right_length >= n1) return 2;   // slice index out of range
        jobs[j] = LevJob{sd.left, sd.right, sd.left_length + 1, sd.right_length + 1, 0, 0, u32(j),
                         (sd.reversed ? 1u : 0u) | (sd.complemented ? 2u : 0u)};
    }
    std::sort(jobs.begin(), jobs.end(), [](const LevJob& a, const LevJob& b) {
        const long double wa = (long double)a.m * a.n, wb = (long double)b.m * b.n;
        return wa != wb ? wa > wb : a.sd_index < b.sd_index;
    });
    DevBuf<u32> d_dist(n_sds, s), d_err(1, s);
    d_err.zero();
    EventTimer timer(s);
    timer.start();
    // batches bounded by the scratch they need (boundary rows: one byte per column and live strip)
    const u64 kScratchCap = u64(4) << 30;
    u64 b0 = 0;
    while (b0 < n_sds) {
        u64 b1 = b0, bytes = 0;
        while (b1 < n_sds) {
            LevJob& jb = jobs[b1];
            const u64 strips = ceil_div(jb.m, u64(kLevStripRows));
            const u64 rows = strips > 1 ? std::min<u64>(strips - 1, kLevWarps) : 0;
            jb.row_stride = ceil_div(jb.n, u64(128)) * 128;
            const u64 need = rows * jb.row_stride;
            if (b1 > b0 && bytes + need > kScratchCap) break;
            jb.bnd_off = bytes;
            bytes += need;
            ++b1;
        }
        DevBuf<u8> d_bnd(bytes + 128, s);
        DevBuf<LevJob> d_jobs(b1 - b0, s);
        CUDA_CHECK(cudaMemcpyAsync(d_jobs.p, jobs.data() + b0, (b1 - b0) * sizeof(LevJob), cudaMemcpyHostToDevice, s));
        levenshtein_kernel<<<unsigned(b1 - b0), kLevWarps * 32, 0, s>>>(d_text, d_jobs.p, d_bnd.p, d_dist.p, d_err.p);
        KERNEL_CHECK();
        count_launch();
        CUDA_CHECK(cudaStreamSynchronize(s));   // jobs[] staging and d_bnd are reused by the next batch
        b0 = b1;
    }
    lev_identity_kernel<<<unsigned(ceil_div(n_sds, 256)), 256, 0, s>>>(d_sds, d_dist.p, n_sds);
    KERNEL_CHECK();
    count_launch();
    timer.stop();
    u32 h_err = 0;
    CUDA_CHECK(cudaMemcpyAsync(&h_err, d_err.p, sizeof h_err, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    if (stats) {
        stats->ms += timer.ms();
        stats->pairs += n_sds;
        for (const LevJob& jb : jobs) stats->cells += jb.m * jb.n;
    }
    return h_err ? 1 : 0;
}

}  // namespace ab200
