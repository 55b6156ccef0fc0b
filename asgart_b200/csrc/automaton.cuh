// automaton.cuh — stage B kernels: processed-index bookkeeping, segment detection, the arm automaton, family
// compaction, and the post-steps FilterNs / ReOrder / ReduceOverlap / Sort (src/bin/asgart.rs:33-96, :481-562).
#pragma once
#include "automaton_core.h"
#include "common.cuh"
#include "scan.cuh"
#include "search.cuh"

namespace ab200 {

// processed index of global probe g = number of processed iterations before it (over all chunks)
__device__ __forceinline__ u64 processed_before(const u32* __restrict__ bits, const u64* __restrict__ wpre, u64 g) {
    const u64 w = g >> 5;
    const u32 r = u32(g & 31u);
    u64 t = wpre[w];
    if (r) t += __popc(bits[w] & ((1u << r) - 1u));
    return t;
}

// per event: chunk, needle-local i, chunk-relative processed index t, segment-head flag
__global__ void event_info_kernel(const u64* __restrict__ ev_probe, u64 n_events, const ChunkDev* __restrict__ chunks,
                                  u32 n_chunks, const u32* __restrict__ bits, const u64* __restrict__ wpre, u32 s, u64 q_ext,
                                  u64* __restrict__ ev_i, u64* __restrict__ ev_t, u32* __restrict__ ev_chunk,
                                  u8* __restrict__ ev_head) {
    const u64 e = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n_events) return;
    const u64 g = ev_probe[e];
    const u32 c = chunk_of_probe(chunks, n_chunks, g);
    const u64 base_t = processed_before(bits, wpre, chunks[c].probe_base);
    const u64 t = processed_before(bits, wpre, g) - base_t;
    bool head = true;
    if (e > 0) {
        const u64 gp = ev_probe[e - 1];
        const u32 cp = chunk_of_probe(chunks, n_chunks, gp);
        if (cp == c) {
            const u64 tp = processed_before(bits, wpre, gp) - base_t;
            head = (t - tp) > q_ext;  // every arm is inactive and flushed before t: hard reset
        }
    }
    ev_i[e] = (g - chunks[c].probe_base + 1) * u64(s);
    ev_t[e] = t;
    ev_chunk[e] = c;
    ev_head[e] = head ? 1 : 0;
}

struct AutoBuffers {
    const u64* ev_i;
    const u64* ev_t;
    const u64* ev_moff;
    const u32* ev_cnt;
    const u32* ev_chunk;
    const u64* seg_first;  // n_segments + 1 entries (last = n_events)
    const u64* matches;
    i64* op_target;
    u64 *a_ls, *a_le, *a_rs, *a_re, *a_death;  // arm store, indexed like matches
    asgart_b200_protosd* out_sd;               // slot array, indexed like matches
    u8* out_flag;                              // 0 empty, 1 duplicon, 3 duplicon that opens a family
    const u64* chunk_tc;                       // processed iterations per chunk
    const ChunkDev* chunks;
};

// v1: one thread per segment (the simulation is sequential in its events; segments are independent)
__global__ void automaton_kernel(AutoBuffers B, AutoParams P, u64 n_segments, u32 reversed_flag, u32 complemented_flag) {
    const u64 sidx = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (sidx >= n_segments) return;
    const u64 e0 = B.seg_first[sidx], e1 = B.seg_first[sidx + 1];
    const u32 c = B.ev_chunk[e0];
    const ChunkDev ch = B.chunks[c];
    const u64 slot0 = B.ev_moff[e0];
    ArmStore arms{B.a_ls + slot0, B.a_le + slot0, B.a_rs + slot0, B.a_re + slot0, B.a_death + slot0};
    u64 cursor = slot0;
    simulate_segment(e0, e1, B.ev_i, B.ev_t, B.ev_moff, B.ev_cnt, B.matches, B.op_target, arms, P, ch.c0, ch.len,
                     B.chunk_tc[c], [&](const SdOut& sd, bool head) {
                         asgart_b200_protosd o;
                         o.left = sd.left; o.right = sd.right;
                         o.left_length = sd.left_length; o.right_length = sd.right_length;
                         o.identity = 0.f;
                         o.reversed = u8(reversed_flag); o.complemented = u8(complemented_flag);
                         o._pad[0] = o._pad[1] = 0;
                         B.out_sd[cursor] = o;
                         B.out_flag[cursor] = head ? 3 : 1;
                         ++cursor;
                     });
}

// ------------------------------------------------------------------------------------------------ post-steps
// FilterNs: count of N over strand[p ..= p+len] for both arms (src/structs.rs:454-467), one warp per duplicon
__global__ void n_content_kernel(const u8* __restrict__ text, const asgart_b200_protosd* __restrict__ sds, u64 n_sds,
                                 u8* __restrict__ keep) {
    const u64 j = (u64(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (j >= n_sds) return;
    const asgart_b200_protosd sd = sds[j];
    u64 cl = 0, cr = 0;
    for (u64 p = sd.left + lane_id(); p <= sd.left + sd.left_length; p += 32) { u8 c = text[p]; cl += (c == 'N' || c == 'n'); }
    for (u64 p = sd.right + lane_id(); p <= sd.right + sd.right_length; p += 32) { u8 c = text[p]; cr += (c == 'N' || c == 'n'); }
    cl = warp_sum_u64(cl);
    cr = warp_sum_u64(cr);
    if (lane_id() == 0) {
        const float l = float(cl) / float(sd.left_length);   // f32 arithmetic like the reference
        const float r = float(cr) / float(sd.right_length);
        keep[j] = fmaxf(l, r) <= 0.2f ? 1 : 0;
    }
}

__device__ __forceinline__ bool subsegment_d(u64 xs, u64 xl, u64 ys, u64 yl) { return xs >= ys && xs + xl <= ys + yl; }
__device__ __forceinline__ bool overlap_d(u64 xs, u64 xl, u64 ys, u64 yl) {
    const u64 xe = xs + xl, ye = ys + yl;
    return (xs >= ys && xs <= ye && xe >= ye) || (ys >= xs && ys <= xe && ye >= xe);
}

// one pass of _reduce (src/bin/asgart.rs:516-551) in place over f[0..n): returns the new length
__device__ u64 reduce_once_d(asgart_b200_protosd* f, u64 n) {
    u64 m = 0;  // `news` = f[0..m); m <= index of the element being inserted, so in-place is safe
    for (u64 r = 0; r < n; ++r) {
        const asgart_b200_protosd x = f[r];
        bool absorbed = false;
        for (u64 w = 0; w < m; ++w) {
            asgart_b200_protosd& y = f[w];
            if (subsegment_d(x.left, x.left_length, y.left, y.left_length) &&
                subsegment_d(x.right, x.right_length, y.right, y.right_length)) { absorbed = true; break; }
            if (subsegment_d(y.left, y.left_length, x.left, x.left_length) &&
                subsegment_d(y.right, y.right_length, x.right, x.right_length)) {
                y.left = x.left; y.right = x.right; y.left_length = x.left_length; y.right_length = x.right_length;
                absorbed = true; break;
            }
            if (overlap_d(x.left, x.left_length, y.left, y.left_length) &&
                overlap_d(x.right, x.right_length, y.right, y.right_length)) {
                // merge (src/bin/asgart.rs:497-513) with its mixed-up lengths (quirk Q5)
                const u64 nl = min(x.left, y.left);
                const u64 ls = max(x.left + x.left_length, y.left + y.right_length) - nl;
                const u64 nr = min(x.right, y.right);
                const u64 rs = max(x.right + x.left_length, y.right + y.right_length) - nr;
                y.left = nl; y.right = nr; y.left_length = ls; y.right_length = rs;
                absorbed = true; break;
            }
        }
        if (!absorbed) { f[m] = x; ++m; }
    }
    return m;
}

// per family (v1: one thread each): drop filtered duplicons, ReOrder, ReduceOverlap to fixpoint, stable Sort by left
__global__ void post_family_kernel(asgart_b200_protosd* __restrict__ sds, const u64* __restrict__ fam_off, u64 n_fam,
                                   const u8* __restrict__ keep, u32 post_mask, u64* __restrict__ new_count) {
    const u64 fidx = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (fidx >= n_fam) return;
    const u64 b = fam_off[fidx], e = fam_off[fidx + 1];
    asgart_b200_protosd* f = sds + b;
    u64 n = e - b;
    if (post_mask & ASGART_B200_POST_FILTER_NS) {
        u64 m = 0;
        for (u64 r = 0; r < n; ++r)
            if (keep[b + r]) { if (m != r) f[m] = f[r]; ++m; }
        n = m;
    }
    if (post_mask & ASGART_B200_POST_REORDER)
        for (u64 r = 0; r < n; ++r)
            if (f[r].left > f[r].right) { const u64 t = f[r].left; f[r].left = f[r].right; f[r].right = t; }  // positions only (Q4)
    if (post_mask & ASGART_B200_POST_REDUCE_OVERLAP) {
        u64 old_size = n;
        u64 cur = reduce_once_d(f, n);
        while (cur < old_size) { old_size = cur; cur = reduce_once_d(f, cur); }
        n = cur;
    }
    if (post_mask & ASGART_B200_POST_SORT) {  // insertion sort is stable, like sort_by
        for (u64 r = 1; r < n; ++r) {
            const asgart_b200_protosd x = f[r];
            u64 w = r;
            while (w > 0 && f[w - 1].left > x.left) { f[w] = f[w - 1]; --w; }
            f[w] = x;
        }
    }
    new_count[fidx] = n;
}

}  // namespace ab200
