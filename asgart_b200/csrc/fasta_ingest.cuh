// fasta_ingest.cuh — FASTA / multiFASTA bytes -> normalised strand, record table and long N-runs, on the device.
//
// Replaces, for a file already in memory, the host-side read_fasta + find_chunks_to_process of prepare_data
// (src/bin/asgart.rs:278-366). The record semantics are those of the bio FASTA reader the reference
// calls (src/bin/asgart.rs:282-290): a line whose first byte is '>' opens a record, every other line is a sequence line
// and contributes its bytes with the trailing white space cut off (`trim_end`); white space INSIDE a sequence line stays
// (and becomes 'N' below, exactly as the reference's normalisation treats any byte outside ATGCN).
//
// Everything is a flag scan (scan.cuh) over the bytes, so the work does not depend on line lengths (a one-line 250 Mbp
// record costs what 4 M lines of 60 columns cost):
//   R   reverse scan, marks: '\n' / non-white byte. A white byte is trailing iff the nearest non-blank event after it is
//       a newline (or the end of the file)                                                   -> class bit per byte
//   H   forward scan, marks: header start ('>' as first byte of a line) / '\n'. A byte lies on a header line iff the
//       latest header start is more recent than the latest newline                            -> keep bit, header bit
//   K   forward scan, counts: kept bytes / header starts. Output pass: compaction + normalisation of the kept bytes
//       (src/bin/asgart.rs:291-301) and one (file offset, strand position) pair per record
//   N   forward scan over the strand, mark: non-N byte. At the last byte of an N-run the mark gives the run's first byte;
//       runs longer than 5000 (src/bin/asgart.rs:326,336) are appended to a list (there are at most n/5001 of them)
// Algorithmic bytes: 1 read per file byte and pass (R, H, K) + 1 write per kept byte + 1 read per strand byte (N).
#pragma once
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "scan.cuh"

namespace ab200 {

constexpr u64 kLongNRun = 5000;   // src/bin/asgart.rs:326

__host__ __device__ __forceinline__ bool fa_space(u8 c) { return c == ' ' || (c >= 9 && c <= 13); }   // ASCII isspace
__host__ __device__ __forceinline__ bool fa_blank(u8 c) { return fa_space(c) && c != '\n'; }

// src/bin/asgart.rs:291-301
__host__ __device__ __forceinline__ u8 fa_normalise(u8 c, bool skip_masked) {
    if (!skip_masked && c >= 'a' && c <= 'z') c = u8(c - 32);
    return (c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N') ? c : u8('N');
}

struct IngestPiece {
    DevBuf<u8> strand;               // kept, normalised bytes of one file (no '$')
    u64 kept = 0;
    std::vector<u64> rec_off;        // file offset of each record's '>'
    std::vector<u64> rec_pos;        // strand position (file-local) at which the record's sequence starts
    std::vector<u64> run_start, run_len;   // maximal N-runs longer than kLongNRun, ascending, file-local strand coordinates
    std::vector<std::string> names;  // record ids (header up to the first white space), filled by the caller
};

enum : u8 { FA_KEEP = 1, FA_HEADER = 2 };

inline void ingest_fasta_device(const u8* d_file, u64 n, bool skip_masked, IngestPiece& out, cudaStream_t stream) {
    using Acc = FlagAcc<u64>;
    out.kept = 0;
    out.rec_off.clear(); out.rec_pos.clear(); out.run_start.clear(); out.run_len.clear();
    if (n == 0) return;
    DevBuf<u8> cls(n, stream);
    u8* d_cls = cls.p;
    {   // R: element 0 stands for the end of the file, element e >= 1 for byte n - e
        auto flags = [=] __device__(u64 e) -> u32 {
            if (e == 0) return 0u;
            const u8 c = d_file[n - e];
            return c == '\n' ? u32(FS_MARK_A) : (fa_space(c) ? 0u : u32(FS_MARK_B));
        };
        auto write = [=] __device__(u64 e, const Acc&, const Acc& inc) {
            if (e == 0) return;
            const u8 c = d_file[n - e];
            d_cls[n - e] = (fa_blank(c) && inc.a >= inc.b) ? 1 : 0;
        };
        FlagScanPlan<u64> plan;
        plan.prepare(flags, n + 1, (Acc*)nullptr, stream);
        plan.finish(flags, write);
    }
    {   // H: element 0 is the (virtual) newline before the file, element e >= 1 is byte e - 1
        auto flags = [=] __device__(u64 e) -> u32 {
            if (e == 0) return 0u;
            const u64 b = e - 1;
            const u8 c = d_file[b];
            if (c == '\n') return u32(FS_MARK_B);
            return (c == '>' && (b == 0 || d_file[b - 1] == '\n')) ? u32(FS_MARK_A) : 0u;
        };
        auto write = [=] __device__(u64 e, const Acc& exc, const Acc& inc) {
            if (e == 0) return;
            const u64 b = e - 1;
            const bool on_header = inc.a > inc.b;
            const bool keep = !on_header && d_file[b] != '\n' && !d_cls[b];
            d_cls[b] = u8((keep ? FA_KEEP : 0) | (inc.a != exc.a ? FA_HEADER : 0));
        };
        FlagScanPlan<u64> plan;
        plan.prepare(flags, n + 1, (Acc*)nullptr, stream);
        plan.finish(flags, write);
    }
    DevBuf<Acc> d_total(1, stream);
    DevBuf<u64> d_rec;
    {   // K
        auto flags = [=] __device__(u64 b) -> u32 {
            const u8 k = d_cls[b];
            return ((k & FA_KEEP) ? u32(FS_CNT_C) : 0u) | ((k & FA_HEADER) ? u32(FS_CNT_D) : 0u);
        };
        FlagScanPlan<u64> plan;
        plan.prepare(flags, n, d_total.p, stream);
        Acc tot(0);
        CUDA_CHECK(cudaMemcpyAsync(&tot, d_total.p, sizeof tot, cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        out.kept = tot.c;
        const u64 records = tot.d;
        out.strand.alloc(out.kept, stream);
        d_rec.alloc(2 * records, stream);
        u8* d_strand = out.strand.p;
        u64* rec = d_rec.p;
        auto write = [=] __device__(u64 b, const Acc& exc, const Acc&) {
            const u8 k = d_cls[b];
            if (k & FA_KEEP) d_strand[exc.c] = fa_normalise(d_file[b], skip_masked);
            if (k & FA_HEADER) { rec[2 * exc.d] = b; rec[2 * exc.d + 1] = exc.c; }
        };
        plan.finish(flags, write);
        std::vector<u64> h(2 * records);
        if (records) CUDA_CHECK(cudaMemcpyAsync(h.data(), rec, h.size() * sizeof(u64), cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        out.rec_off.resize(records); out.rec_pos.resize(records);
        for (u64 r = 0; r < records; ++r) { out.rec_off[r] = h[2 * r]; out.rec_pos[r] = h[2 * r + 1]; }
    }
    const u64 m = out.kept;
    if (m > kLongNRun) {   // N: element 0 is a virtual non-N byte before the strand, element e >= 1 is strand byte e - 1
        const u8* d_strand = out.strand.p;
        const u64 cap = m / (kLongNRun + 1) + 1;
        DevBuf<u64> d_runs(2 * cap + 1, stream);
        CUDA_CHECK(cudaMemsetAsync(d_runs.p, 0, sizeof(u64), stream));
        u64* cnt = d_runs.p;
        u64* runs = d_runs.p + 1;
        auto flags = [=] __device__(u64 e) -> u32 {
            if (e == 0) return 0u;
            return d_strand[e - 1] != 'N' ? u32(FS_MARK_A) : 0u;
        };
        auto write = [=] __device__(u64 e, const Acc&, const Acc& inc) {
            if (e == 0 || d_strand[e - 1] != 'N') return;
            if (e != m && d_strand[e] == 'N') return;             // not the last byte of its run
            const u64 first = inc.a;                             // element a = byte a - 1 is the last non-N one
            const u64 len = e - first;
            if (len > kLongNRun) {
                const u64 slot = atomicAdd(reinterpret_cast<unsigned long long*>(cnt), 1ull);
                if (slot < cap) { runs[2 * slot] = first; runs[2 * slot + 1] = len; }
            }
        };
        FlagScanPlan<u64> plan;
        plan.prepare(flags, m + 1, (Acc*)nullptr, stream);
        plan.finish(flags, write);
        u64 h_cnt = 0;
        CUDA_CHECK(cudaMemcpyAsync(&h_cnt, cnt, sizeof h_cnt, cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        h_cnt = std::min(h_cnt, cap);
        std::vector<u64> h(2 * h_cnt);
        if (h_cnt) CUDA_CHECK(cudaMemcpyAsync(h.data(), runs, h.size() * sizeof(u64), cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        std::vector<std::pair<u64, u64>> v(h_cnt);
        for (u64 i = 0; i < h_cnt; ++i) v[i] = {h[2 * i], h[2 * i + 1]};
        std::sort(v.begin(), v.end());
        for (auto& p : v) { out.run_start.push_back(p.first); out.run_len.push_back(p.second); }
    }
}

// chunks_to_process of one file (src/bin/asgart.rs:317-366 applied per fragment, :381-387): inside each fragment the
// maximal regions between N-runs longer than kLongNRun; a run that crosses a fragment border counts on each side with the
// part that lies there. `off` = strand position of the file's first base.
inline void ingest_chunks(const IngestPiece& p, u64 off, std::vector<asgart_b200_chunk>& chunks,
                          std::vector<u64>& frag_pos, std::vector<u64>& frag_len) {
    size_t ri = 0;
    const size_t R = p.rec_pos.size();
    for (size_t r = 0; r < R; ++r) {
        const u64 fs = p.rec_pos[r], fe = r + 1 < R ? p.rec_pos[r + 1] : p.kept;
        frag_pos.push_back(off + fs);
        frag_len.push_back(fe - fs);
        const size_t first = chunks.size();
        u64 cur = fs;
        while (ri < p.run_start.size() && p.run_start[ri] + p.run_len[ri] <= fs) ++ri;
        for (size_t j = ri; j < p.run_start.size() && p.run_start[j] < fe; ++j) {
            const u64 ps = std::max(p.run_start[j], fs), pe = std::min(p.run_start[j] + p.run_len[j], fe);
            if (pe - ps <= kLongNRun) continue;
            if (ps > cur) chunks.push_back({off + cur, ps - cur});
            cur = pe;
        }
        if (fe > cur) chunks.push_back({off + cur, fe - cur});
        if (chunks.size() == first) chunks.push_back({off + fs, fe - fs});
    }
}

}  // namespace ab200
