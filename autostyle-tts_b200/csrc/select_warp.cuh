// Warp-per-query level select (K4), used inside the persistent tensor-core scan kernel: after a level has been
// scanned by the whole grid, every warp of the grid takes queries in turn and
//   intermediate level: finds the rank-j key of what the query has collected (the next level's threshold) and
//                       compacts the survivors to the front of the query's candidate buffer;
//   final level:        hands the best K' keys to the rescoring stage together with `bound`, an upper bound on the
//                       scan score of every row that is NOT among them.
// Same semantics as select_level_kernel (search.cu), which stays for the warp-dot (gemv) path.
//   <= 256 keys  : bitonic sort in registers (8 keys per lane)
//   dense level  : pivot method - the rank-j value of the 32 lane maxima is a lower bound of the rank-j key; the few
//                  keys above it are compacted into the warp's shared-memory list and sorted in registers
//   otherwise    : 8-pass byte-wise radix select (warp histogram in shared memory) + ordered in-place compaction
#pragma once

#include "avs_internal.h"

#define AVS_ST_OVERFLOW 1

// The select reads its arguments straight from the kernel's AvsScanPlan (a __grid_constant__ parameter: no copy on
// the stack); n_eff = rows that may be returned (filter-allowed count).
typedef AvsScanPlan WarpSelectArgs;

template <int EPL>
__device__ __forceinline__ void wsel_sort_desc(u64 (&k)[EPL], int lane) {
#pragma unroll
    for (int k2 = 2; k2 <= 32 * EPL; k2 <<= 1) {
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int jr = j >> 5;
#pragma unroll
                for (int r = 0; r < EPL; ++r) {
                    const int pr = r ^ jr;
                    if (pr > r) {
                        const bool desc = (((32 * r) & k2) == 0);
                        const u64 x = k[r], y = k[pr];
                        const bool sw = desc ? (x < y) : (x > y);
                        k[r] = sw ? y : x;
                        k[pr] = sw ? x : y;
                    }
                }
            } else {
                const bool lower = (lane & j) == 0;
#pragma unroll
                for (int r = 0; r < EPL; ++r) {
                    const bool desc = (((lane + 32 * r) & k2) == 0);
                    const u64 other = __shfl_xor_sync(0xffffffffu, k[r], j);
                    const u64 mx = k[r] > other ? k[r] : other, mn = k[r] > other ? other : k[r];
                    k[r] = (desc == lower) ? mx : mn;
                }
            }
        }
    }
}

// element e (0-based rank) of a register-sorted array: lane e % 32, register e / 32
template <int EPL>
__device__ __forceinline__ u64 wsel_pick(const u64 (&k)[EPL], int e) {
    u64 v = 0;
#pragma unroll
    for (int r = 0; r < EPL; ++r) v = (r == (e >> 5)) ? k[r] : v;
    return __shfl_sync(0xffffffffu, v, e & 31);
}

// the final level's `bound` (see select_emit_scalar in search.cu); one lane calls it
__device__ __forceinline__ void wsel_emit_final(const WarpSelectArgs& a, int q, int n, u64 key_kp, bool lost) {
    const int m = n < a.kprime ? n : a.kprime;
    a.topn[q] = m;
    float b;
    int st = 0;
    if (lost || (a.status[q] & AVS_ST_OVERFLOW)) { b = INFINITY; st = AVS_ST_OVERFLOW; }   // entries were lost
    else if (n > a.kprime) b = avs_key_score(key_kp);                  // rows outside the K' candidates <= K'-th key
    else if ((int64_t)n >= a.n_eff) b = -INFINITY;                    // every row is a candidate
    else b = a.tau[q] == 0ull ? -INFINITY : avs_key_score(a.tau[q]);   // rows outside < threshold
    a.bound[q] = b;
    a.status[q] |= st;
}

// Select among n_src <= 32 * EPL keys held in registers (key e in lane e % 32, register e / 32).  Every key's rank is
// the number of keys above it (keys are distinct: the row index is part of the key; empty slots hold 0 and are skipped),
// found by broadcasting the keys one by one: a short real loop instead of an unrolled sorting network - a lone warp
// selecting for a single query runs ~10 instructions per key, not the ~6 k of a 256-key bitonic sort.  Writing key ->
// slot[rank] leaves the survivors sorted.
// Output phase of the small selects: k[] = the keys (key e in lane e % 32, register e / 32), rk[] = their ranks.
template <int EPL>
__device__ __forceinline__ void wsel_emit(const WarpSelectArgs& a, int q, int lane, const u64 (&k)[EPL], const int (&rk)[EPL], int n_real,
                                          int jj, bool is_final, int dense_total, int k_eps, u64* c, int total_in, bool lost) {
    // key of a given rank (0 when no such key): the one lane that holds it publishes it
    auto key_of_rank = [&](int e) -> u64 {
        u64 v = 0;
#pragma unroll
        for (int r = 0; r < EPL; ++r) v = (k[r] != 0ull && rk[r] == e) ? k[r] : v;
        const unsigned who = __ballot_sync(0xffffffffu, v != 0ull);
        return who ? __shfl_sync(0xffffffffu, v, __ffs(who) - 1) : 0ull;
    };
    if (!is_final) {
        const u64 tau_old = a.tau[q];                                   // every lane reads it before lane 0 updates it below
        __syncwarp();
        int keep = n_real >= jj ? jj : n_real;
        u64 tau_new = n_real >= jj ? key_of_rank(jj - 1) : tau_old;
        // Last threshold (eps rule, see select_level_kernel): at least 2.5 eps below the k-th scan score seen so far
        if (k_eps > 0 && n_real >= k_eps && !lost) {
            const u64 kk = key_of_rank(k_eps - 1);
            u64 t_eps = avs_make_key(avs_key_score(kk) - 2.5f * a.eps[q], 0xFFFFFFFFu);
            // never below the threshold in force while these keys were collected: rows under THAT one were dropped
            // already, and the final bound ("every row outside the buffer scores below tau") must hold for them too
            if (t_eps < tau_old) t_eps = tau_old;
            if (t_eps < tau_new) {
                tau_new = t_eps;
                int above = 0;
#pragma unroll
                for (int r = 0; r < EPL; ++r) above += (k[r] >= tau_new && k[r] != 0ull) ? 1 : 0;
#pragma unroll
                for (int o = 16; o; o >>= 1) above += __shfl_xor_sync(0xffffffffu, above, o);
                keep = above;
            }
        }
#pragma unroll
        for (int r = 0; r < EPL; ++r)
            if (k[r] != 0ull && rk[r] < keep) c[rk[r]] = k[r];
        if (lane == 0) {
            a.tau[q] = tau_new;
            a.cnt[q] = keep;
            if (lost) a.status[q] |= AVS_ST_OVERFLOW;
        }
    } else {
        const int m = n_real < a.kprime ? n_real : a.kprime;
        u64* tk = a.topkeys + (size_t)q * a.kprime;
#pragma unroll
        for (int r = 0; r < EPL; ++r) {
            if (k[r] != 0ull) {
                if (rk[r] < m) tk[rk[r]] = k[r];
                c[rk[r]] = k[r];                                      // sorted and compacted: the wide-rescoring stage reads it
            }
        }
        for (int i = m + lane; i < a.kprime; i += 32) tk[i] = 0ull;
        const u64 key_kp = n_real >= a.kprime ? key_of_rank(a.kprime - 1) : 0ull;
        if (lane == 0) {
            a.cnt[q] = dense_total > 0 ? n_real : total_in;           // dense level: compacted, no empty slots left in front
            wsel_emit_final(a, q, n_real, key_kp, lost);
        }
    }
    __syncwarp();
}

// Select among n_src <= 32 * EPL keys held in registers (key e in lane e % 32, register e / 32).  Every key's rank is
// the number of keys above it (keys are distinct: the row index is part of the key; empty slots hold 0 and are skipped),
// found by broadcasting the keys one by one: a short real loop instead of an unrolled sorting network - a lone warp
// selecting for a single query runs ~10 instructions per key, not the ~6 k of a 256-key bitonic sort.  Writing key ->
// slot[rank] leaves the survivors sorted.
template <int EPL>
__device__ __forceinline__ void wsel_small(const WarpSelectArgs& a, int q, int lane, int jj, bool is_final, int dense_total,
                                           int k_eps, const u64* src, int n_src, u64* c, int total_in, bool lost) {
    u64 k[EPL];
    int rk[EPL];
    int nz = 0;
#pragma unroll
    for (int r = 0; r < EPL; ++r) {
        const int i = lane + 32 * r;
        k[r] = i < n_src ? src[i] : 0ull;
        rk[r] = 0;
        nz += k[r] != 0ull;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, o);
#pragma unroll
    for (int r = 0; r < EPL; ++r) {
        const int lim = n_src - 32 * r < 32 ? n_src - 32 * r : 32;      // uniform across the warp
        for (int l = 0; l < lim; ++l) {
            const u64 b = __shfl_sync(0xffffffffu, k[r], l);
#pragma unroll
            for (int r2 = 0; r2 < EPL; ++r2) rk[r2] += b > k[r2] ? 1 : 0;
        }
    }
    __syncwarp();                                                      // every key has been read: `src` may alias `c`
    wsel_emit<EPL>(a, q, lane, k, rk, nz, jj, is_final, dense_total, k_eps, c, total_in, lost);
}

// `list`: 256 u64 of shared memory owned by this warp (also used as a 256-bin int histogram by the radix path).
__device__ __noinline__ void warp_select_level(const WarpSelectArgs& a, int q, int lane, int j_rank, bool is_final,
                                               int dense_total, int k_eps, u64* list, bool boot_lists = false) {
    u64* c = a.cand + (size_t)q * a.cap;
    const int total_in = dense_total > 0 ? dense_total : a.cnt[q];
    const bool lost = dense_total == 0 && total_in > a.cap;       // more keys were offered than the buffer holds
    const int n = total_in < a.cap ? total_in : a.cap;            // slots to look at
    int jj = j_rank;
    if (lost) { jj = (int)(((long long)j_rank * a.cap) / total_in); if (jj < 1) jj = 1; }

    // ---- dense (threshold-free) level with a small rank: pivot method, two passes over the keys ----
    // the rank-j value of the 32 lane maxima is a lower bound of the rank-j key; the keys above it (j .. a few dozen) are
    // compacted into the warp's list and then take the register-sort path below
    const u64* src = c;
    int n_src = n;
    // ---- boot level (lists of 8 keys, the best one in slot 0): pivot from the list heads alone ----
    // the rank-j value of the 32 lane maxima over the HEADS is a lower bound of the rank-j key; a key at or above it sits
    // in a list whose head is at or above it - a handful of lists.  One pass over an eighth of the keys, then those lists.
    bool heads_done = false;
    if (boot_lists && n > 256 && !is_final && jj <= 32 && k_eps == 0) {
        const int n_lists = n >> 3;
        u64 lm = 0;
#pragma unroll 8
        for (int l = lane; l < n_lists; l += 32) { const u64 key = c[8 * l]; lm = key > lm ? key : lm; }
        u64 k1[1] = {lm};
        wsel_sort_desc<1>(k1, lane);
        const u64 P = __shfl_sync(0xffffffffu, k1[0], jj - 1);
        if (P != 0ull) {
            int m = 0;
            for (int l0 = 0; l0 < n_lists; l0 += 32) {
                const int l = l0 + lane;
                const bool hot = l < n_lists && c[8 * l] >= P;
                if (__any_sync(0xffffffffu, hot)) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const u64 key = hot ? c[8 * l + i] : 0ull;
                        const bool in = key >= P;
                        const unsigned bal = __ballot_sync(0xffffffffu, in);
                        const int pos = m + __popc(bal & ((1u << lane) - 1));
                        if (in && pos < 256) list[pos] = key;
                        m += __popc(bal);
                    }
                }
            }
            __syncwarp();
            if (m <= 256) { src = list; n_src = m; heads_done = true; }   // m >= jj by construction
        }
    }
    if (heads_done) {
        // fall through to the register select on `list`
    } else
    // (also for a thresholded level that collected more than 256 keys - the tail of the survivor-count distribution: one
    // such query among a thousand would otherwise send its warp through the 8-pass radix select while the grid waits)
    if (n > 256 && !lost && !is_final && jj <= 32 && k_eps == 0) {   // eps rule: the threshold may end up below the pivot
        u64 lm = 0;
#pragma unroll 8
        for (int i = lane; i < n; i += 32) { const u64 key = c[i]; lm = key > lm ? key : lm; }   // 8 loads in flight per lane
        u64 k1[1] = {lm};
        wsel_sort_desc<1>(k1, lane);
        const u64 P = __shfl_sync(0xffffffffu, k1[0], jj - 1);   // j distinct keys are >= P
        if (P != 0ull) {
            int m = 0;
#pragma unroll 8
            for (int i0 = 0; i0 < n; i0 += 32) {
                const int i = i0 + lane;
                const u64 key = i < n ? c[i] : 0ull;
                const bool in = key >= P;
                const unsigned bal = __ballot_sync(0xffffffffu, in);
                const int pos = m + __popc(bal & ((1u << lane) - 1));
                if (in && pos < 256) list[pos] = key;
                m += __popc(bal);
            }
            __syncwarp();
            if (m <= 256) { src = list; n_src = m; }              // m >= jj by construction
        }
        // else: few real keys (row filter) or a pile-up at the pivot -> radix select below
    }

    if (n_src <= 256) {
        // ---- up to 8 keys per lane in registers: rank every key by counting ----
        // the rank loop costs n_src * (2 + 3 * EPL) instructions: registers per lane follow the key count closely
        if (n_src <= 32) wsel_small<1>(a, q, lane, jj, is_final, dense_total, k_eps, src, n_src, c, total_in, lost);
        else if (n_src <= 64) wsel_small<2>(a, q, lane, jj, is_final, dense_total, k_eps, src, n_src, c, total_in, lost);
        else if (n_src <= 96) wsel_small<3>(a, q, lane, jj, is_final, dense_total, k_eps, src, n_src, c, total_in, lost);
        else if (n_src <= 128) wsel_small<4>(a, q, lane, jj, is_final, dense_total, k_eps, src, n_src, c, total_in, lost);
        else if (n_src <= 192) wsel_small<6>(a, q, lane, jj, is_final, dense_total, k_eps, src, n_src, c, total_in, lost);
        else wsel_small<8>(a, q, lane, jj, is_final, dense_total, k_eps, src, n_src, c, total_in, lost);
        return;
    }

    // ---- general case: radix select of the rank-`want` key, then a compaction ----
    int* hist = reinterpret_cast<int*>(list);
    int nzc = 0;
    for (int i = lane; i < n; i += 32) nzc += c[i] != 0ull;
#pragma unroll
    for (int o = 16; o; o >>= 1) nzc += __shfl_xor_sync(0xffffffffu, nzc, o);
    const int n_real = nzc;
    const int want = is_final ? (n_real < a.kprime ? n_real : a.kprime) : (n_real >= jj ? jj : n_real);
    // the key of rank `rank0` (0-based, descending) among the real keys of c[0, n): 8 byte-wise histogram passes
    auto radix_key = [&](int rank0) -> u64 {
        u64 prefix = 0, maskb = 0;
        int rank = rank0;
        for (int byte = 7; byte >= 0; --byte) {
#pragma unroll
            for (int b = 0; b < 8; ++b) hist[lane * 8 + b] = 0;
            __syncwarp();
            for (int i0 = 0; i0 < n; i0 += 32) {
                const int i = i0 + lane;
                const u64 key = i < n ? c[i] : 0ull;
                const bool in = i < n && key != 0ull && (key & maskb) == prefix;
                const int bin = in ? (int)((key >> (8 * byte)) & 0xFFull) : 256 + lane;   // non-members: unique dummies
                const unsigned peers = __match_any_sync(0xffffffffu, bin);
                if (in && lane == __ffs(peers) - 1) hist[bin] += __popc(peers);          // one lane per distinct bin: no conflict
                __syncwarp();
            }
            // lane l owns bins 255-8l .. 248-8l (descending key order)
            int loc[8], sum = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) { loc[b] = hist[255 - 8 * lane - b]; sum += loc[b]; }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t2 = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t2; }
            int accb = incl - sum;
            int sel_bin = -1, sel_acc = 0;
            if (rank >= accb && rank < incl) {
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    if (sel_bin < 0 && rank < accb + loc[b]) { sel_bin = 255 - 8 * lane - b; sel_acc = accb; }
                    accb += loc[b];
                }
            }
            const unsigned owner = __ballot_sync(0xffffffffu, sel_bin >= 0);
            const int src = __ffs(owner) - 1;
            sel_bin = __shfl_sync(0xffffffffu, sel_bin, src);
            sel_acc = __shfl_sync(0xffffffffu, sel_acc, src);
            prefix |= (u64)sel_bin << (8 * byte);
            maskb |= 0xFFull << (8 * byte);
            rank -= sel_acc;
            __syncwarp();
        }
        return prefix;
    };
    const u64 Pk = want > 0 ? radix_key(want - 1) : ~0ull;         // the rank-(want-1) key; exactly `want` keys are >= it
    if (!is_final) {
        // Last threshold under the eps rule (see wsel_emit; round 2 found this path ignoring it: levels that collect more
        // than 256 keys - half the queries of a 2.5 M x 3072 shard - kept the rank-j threshold, and ~1 % of them then
        // failed the wide-rescoring certificate and paid for an exact scan of the fp32 master in EVERY search)
        u64 thr = Pk;
        bool lowered = false;
        if (k_eps > 0 && n_real >= k_eps && !lost) {
            const u64 kk = k_eps == want ? Pk : radix_key(k_eps - 1);
            u64 t_eps = avs_make_key(avs_key_score(kk) - 2.5f * a.eps[q], 0xFFFFFFFFu);
            const u64 tau_old = a.tau[q];                          // (lane 0 writes it only after the compaction's warp syncs)
            if (t_eps < tau_old) t_eps = tau_old;                  // rows under the threshold in force were dropped already
            const u64 tau_cur = n_real >= jj ? Pk : tau_old;
            if (t_eps < tau_cur) { thr = t_eps; lowered = true; }
        }
        // ordered in-place compaction: the write cursor never passes the read cursor
        int m = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const u64 key = i < n ? c[i] : 0ull;
            const bool in = key >= thr && key != 0ull;
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            __syncwarp();
            if (in) c[m + __popc(bal & ((1u << lane) - 1))] = key;
            m += __popc(bal);
            __syncwarp();
        }
        if (lane == 0) {
            if (n_real >= jj || lowered) a.tau[q] = thr;
            a.cnt[q] = m;                                          // == want unless the eps rule lowered the threshold
            if (lost) a.status[q] |= AVS_ST_OVERFLOW;
        }
    } else {
        int m = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const u64 key = i < n ? c[i] : 0ull;
            const bool in = key >= Pk && key != 0ull;
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (in) a.topkeys[(size_t)q * a.kprime + m + __popc(bal & ((1u << lane) - 1))] = key;
            m += __popc(bal);
        }
        for (int i = want + lane; i < a.kprime; i += 32) a.topkeys[(size_t)q * a.kprime + i] = 0ull;
        if (lane == 0) {
            a.cnt[q] = dense_total > 0 ? n : total_in;             // dense level: slots (zero = empty), not keys
            wsel_emit_final(a, q, n_real, Pk, lost);
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Lean CTA select for the common small-batch cases (one query per CTA, 256 threads): a chain of as few dependent steps
// as possible, because the whole grid waits for it.
//   threshold-free level (dense or boot lists, <= 2048 slots, rank <= 32): every thread loads 8 slots at once and
//       publishes their maximum; the rank-j maximum (counted against the 256 maxima in shared memory, broadcast reads)
//       is a lower bound of the rank-j key; the few keys at or above it are compacted into shared memory;
//   <= 512 keys: every thread ranks its (at most two) keys by counting against all keys in shared memory; writing
//       key -> slot[rank] leaves them sorted.
// Same outputs as cta_select_level.  Returns false (uniformly, before writing anything) when the case is not covered.
//   `scratch`: >= 1300 u64 of shared memory; `tr`: optional phase stamps (profiling).
// ---------------------------------------------------------------------------------------------
template <typename SyncFn>
__device__ __noinline__ bool cta_select_fast(const WarpSelectArgs& a, int q, int tid, int j_rank, bool is_final,
                                             int dense_total, int k_eps, u64* scratch, SyncFn sync, u64* tr) {
    u64* const maxima = scratch;                                     // [256]
    u64* const list = scratch + 256;                                 // [1024]
    u64* const pub = scratch + 1280;                                 // [0] rank-(j-1) key, [1] rank-(k_eps-1) key, [2] K'-th key, [3] pivot
    int* const misc = reinterpret_cast<int*>(scratch + 1284);        // [0] list fill, [1] real keys, [2] keys >= eps threshold
    auto stamp = [&](int i) {
        if (tr != nullptr && tid == 0) { u64 t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); tr[i] = t; }
    };
    stamp(0);
    u64* c = a.cand + (size_t)q * a.cap;
    const int total_in = dense_total > 0 ? dense_total : a.cnt[q];
    if (dense_total == 0 && total_in > a.cap) return false;          // keys were lost: the general path scales the rank
    const u64 tau_old = a.tau[q];                                    // read by every thread BEFORE thread 0 may update it (below)
    const int n = total_in;
    // (under the eps rule the threshold may end up BELOW the pivot, whose compaction would already have dropped rows)
    const bool pivot_case = dense_total > 0 && !is_final && j_rank <= 32 && n > 512 && n <= 2048 && k_eps == 0;
    if (!pivot_case && n > 512) return false;
    const int lane = tid & 31;
    if (tid < 4) { pub[tid] = 0ull; misc[tid] = 0; }
    int n_src = n;
    if (pivot_case) {
        u64 mine[8];
        u64 lm = 0;
        {
            const ulonglong2* p = reinterpret_cast<const ulonglong2*>(c + 8 * tid);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                ulonglong2 v = make_ulonglong2(0ull, 0ull);
                if (8 * tid + 2 * r < n) v = p[r];                   // n is a multiple of 8 for both kinds of level
                mine[2 * r] = v.x; mine[2 * r + 1] = v.y;
                lm = v.x > lm ? v.x : lm;
                lm = v.y > lm ? v.y : lm;
            }
        }
        maxima[tid] = lm;
        sync();
        stamp(1);
        int rk = 0;
#pragma unroll 8
        for (int i = 0; i < 256; ++i) rk += maxima[i] > lm ? 1 : 0;
        if (lm != 0ull && rk == j_rank - 1) pub[3] = lm;             // keys are distinct: at most one thread
        sync();
        stamp(2);
        const u64 P = pub[3];
        if (P == 0ull) return false;                                 // fewer than j real maxima (row filter): general path
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (mine[r] >= P) { const int pos = atomicAdd(&misc[0], 1); list[pos] = mine[r]; }   // <= 8 j <= 256 keys
        sync();
        stamp(3);
        n_src = misc[0];
    } else {
        for (int i = tid; i < n; i += 256) list[i] = c[i];
        sync();
        stamp(3);
    }
    // ---- rank every key by counting (n_src <= 512: two keys per thread) ----
    const u64 k0 = tid < n_src ? list[tid] : 0ull, k1 = tid + 256 < n_src ? list[tid + 256] : 0ull;
    int r0 = 0, r1 = 0;
    if (n_src <= 256) {
#pragma unroll 8
        for (int i = 0; i < n_src; ++i) r0 += list[i] > k0 ? 1 : 0;
    } else {
#pragma unroll 8
        for (int i = 0; i < n_src; ++i) { const u64 b = list[i]; r0 += b > k0 ? 1 : 0; r1 += b > k1 ? 1 : 0; }
    }
    {
        int real = (k0 != 0ull) + (k1 != 0ull);
#pragma unroll
        for (int o = 16; o; o >>= 1) real += __shfl_xor_sync(0xffffffffu, real, o);
        if (lane == 0 && real) atomicAdd(&misc[1], real);
    }
    const int jj = j_rank;
    if (!is_final) {
        if (k0 != 0ull && r0 == jj - 1) pub[0] = k0;
        if (k1 != 0ull && r1 == jj - 1) pub[0] = k1;
        if (k_eps > 0) {
            if (k0 != 0ull && r0 == k_eps - 1) pub[1] = k0;
            if (k1 != 0ull && r1 == k_eps - 1) pub[1] = k1;
        }
    } else {
        if (k0 != 0ull && r0 == a.kprime - 1) pub[2] = k0;
        if (k1 != 0ull && r1 == a.kprime - 1) pub[2] = k1;
    }
    sync();                                                          // ranks done: `list` has been read by everybody
    stamp(4);
    const int n_real = misc[1];
    if (!is_final) {
        int keep = n_real >= jj ? jj : n_real;
        u64 tau_new = n_real >= jj ? pub[0] : tau_old;
        bool lowered = false;
        // Last threshold (eps rule, see select_level_kernel): at least 2.5 eps below the k-th scan score seen so far
        if (k_eps > 0 && n_real >= k_eps) {
            u64 t_eps = avs_make_key(avs_key_score(pub[1]) - 2.5f * a.eps[q], 0xFFFFFFFFu);
            if (t_eps < tau_old) t_eps = tau_old;                   // see wsel_emit: never below the threshold in force
            if (t_eps < tau_new) { tau_new = t_eps; lowered = true; }
        }
        if (lowered) {                                               // uniform: every thread read the same shared values
            int above = (k0 != 0ull && k0 >= tau_new) + (k1 != 0ull && k1 >= tau_new);
#pragma unroll
            for (int o = 16; o; o >>= 1) above += __shfl_xor_sync(0xffffffffu, above, o);
            if (lane == 0 && above) atomicAdd(&misc[2], above);
            sync();
            keep = misc[2];
        }
        if (k0 != 0ull && r0 < keep) c[r0] = k0;
        if (k1 != 0ull && r1 < keep) c[r1] = k1;
        if (tid == 0) { a.tau[q] = tau_new; a.cnt[q] = keep; }
    } else {
        const int m = n_real < a.kprime ? n_real : a.kprime;
        u64* tk = a.topkeys + (size_t)q * a.kprime;
        if (k0 != 0ull) { if (r0 < m) tk[r0] = k0; c[r0] = k0; }    // sorted and compacted: the wide-rescoring stage reads it
        if (k1 != 0ull) { if (r1 < m) tk[r1] = k1; c[r1] = k1; }
        for (int i = m + tid; i < a.kprime; i += 256) tk[i] = 0ull;
        if (tid == 0) {
            a.cnt[q] = dense_total > 0 ? n_real : total_in;           // dense level: compacted, no empty slots left in front
            wsel_emit_final(a, q, n_real, n_real >= a.kprime ? pub[2] : 0ull, false);
        }
    }
    sync();
    stamp(5);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Few queries (at most one per CTA - the reference's batch-1 search): the 8 epilogue warps of a CTA select for ONE query
// together instead of leaving it to a lone warp whose latency the whole grid waits for.  Same results as
// warp_select_level; the rank counting (the n^2 part) and the dense level's slot scan are split 8 ways.
//   `scratch`: >= 5.5 KB of shared memory of the CTA, idle during the select; sync() = barrier of the 8 warps.
// Every one of the 8 warps must call this (uniform control flow: the barriers are inside).
// ---------------------------------------------------------------------------------------------
template <typename SyncFn>
__device__ __noinline__ void cta_select_level(const WarpSelectArgs& a, int q, int w, int lane, int j_rank, bool is_final,
                                              int dense_total, int k_eps, u64* scratch, SyncFn sync) {
    u64* const list = scratch;                                       // [256] compacted keys of the dense level
    u64* const maxima = scratch + 256;                               // [256] lane maxima of the 8 warps
    int* const ranks = reinterpret_cast<int*>(scratch + 512);        // [256]
    int* const misc = ranks + 256;                                   // [0] list fill, [1] fallback flag
    u64* const pivot = scratch + 512 + 160;                          // [1]
    u64* c = a.cand + (size_t)q * a.cap;
    const int total_in = dense_total > 0 ? dense_total : a.cnt[q];
    const bool lost = dense_total == 0 && total_in > a.cap;
    const int n = total_in < a.cap ? total_in : a.cap;
    int jj = j_rank;
    if (lost) { jj = (int)(((long long)j_rank * a.cap) / total_in); if (jj < 1) jj = 1; }
    const u64* src = c;
    int n_src = n;
    const int tid = w * 32 + lane;
    if (tid < 256) ranks[tid] = 0;
    if (tid == 0) { misc[0] = 0; misc[1] = 0; *pivot = 0ull; }
    sync();

    // ---- dense level, small rank: pivot = rank-j value of the 256 lane maxima, each warp scanning an eighth of the slots ----
    if (n > 256 && n <= 2048 && dense_total > 0 && !is_final && jj <= 32 && k_eps == 0) {
        u64 mine[8];
        u64 lm = 0;
        const int per = (n + 7) >> 3, lo = w * per, hi = lo + per < n ? lo + per : n;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = lo + lane + 32 * r;
            mine[r] = i < hi ? c[i] : 0ull;
            lm = mine[r] > lm ? mine[r] : lm;
        }
        maxima[tid] = lm;
        sync();
        u64 k[8];
        int rk[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) { k[r] = maxima[lane + 32 * r]; rk[r] = 0; }
        {   // this warp counts against the maxima of register w
            u64 kw = 0;
#pragma unroll
            for (int r = 0; r < 8; ++r) kw = r == w ? k[r] : kw;
            for (int l = 0; l < 32; ++l) {
                const u64 b = __shfl_sync(0xffffffffu, kw, l);
#pragma unroll
                for (int r2 = 0; r2 < 8; ++r2) rk[r2] += b > k[r2] ? 1 : 0;
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) if (rk[r]) atomicAdd(&ranks[lane + 32 * r], rk[r]);
        sync();
        // equal maxima (several lanes without a key: 0) share a rank; the rank-(jj-1) one, if it is a real key, is the pivot
        if (w == 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r) if (k[r] != 0ull && ranks[lane + 32 * r] == jj - 1) *pivot = k[r];
        }
        sync();
        const u64 P = *pivot;
        if (P != 0ull) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (mine[r] >= P) { const int pos = atomicAdd(&misc[0], 1); if (pos < 256) list[pos] = mine[r]; }
        }
        if (tid < 256) ranks[tid] = 0;
        sync();
        const int m = misc[0];
        if (P != 0ull && m <= 256) { src = list; n_src = m; }         // m >= jj by construction
    }

    if (n_src <= 256) {
        u64 k[8];
        int rk[8];
        int nz = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = lane + 32 * r;
            k[r] = i < n_src ? src[i] : 0ull;
            rk[r] = 0;
            nz += k[r] != 0ull;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, o);
        const int lim = n_src - 32 * w < 32 ? n_src - 32 * w : 32;   // this warp broadcasts the keys of register w
        if (lim > 0) {
            u64 kw = 0;
#pragma unroll
            for (int r = 0; r < 8; ++r) kw = r == w ? k[r] : kw;
            for (int l = 0; l < lim; ++l) {
                const u64 b = __shfl_sync(0xffffffffu, kw, l);
#pragma unroll
                for (int r2 = 0; r2 < 8; ++r2) rk[r2] += b > k[r2] ? 1 : 0;
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) if (rk[r]) atomicAdd(&ranks[lane + 32 * r], rk[r]);
        }
        sync();                                                       // all partial ranks are in; every warp has read `src`
        if (w == 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r) rk[r] = ranks[lane + 32 * r];
            wsel_emit<8>(a, q, lane, k, rk, nz, jj, is_final, dense_total, k_eps, c, total_in, lost);
        }
    } else {
        // ---- large sets: 8-pass byte-wise radix select, the 8 warps histogramming an eighth of the keys each ----
        int* const hist = ranks;                                      // 256 bins
        {
            int nzc = 0;
            for (int i = tid; i < n; i += 256) nzc += c[i] != 0ull;
#pragma unroll
            for (int o = 16; o; o >>= 1) nzc += __shfl_xor_sync(0xffffffffu, nzc, o);
            if (lane == 0 && nzc) atomicAdd(&misc[1], nzc);
        }
        sync();
        const int n_real = misc[1];
        const int want = is_final ? (n_real < a.kprime ? n_real : a.kprime) : (n_real >= jj ? jj : n_real);
        u64 prefix = 0, maskb = 0;
        int rank = want - 1;
        if (want > 0) {
            for (int byte = 7; byte >= 0; --byte) {
                if (tid < 256) hist[tid] = 0;
                sync();
                for (int i0 = 0; i0 < n; i0 += 256) {                 // whole warps stay converged for the match below
                    const int i = i0 + tid;
                    const u64 key = i < n ? c[i] : 0ull;
                    const bool in = i < n && key != 0ull && (key & maskb) == prefix;
                    const int bin = in ? (int)((key >> (8 * byte)) & 0xFFull) : 256 + lane;
                    const unsigned peers = __match_any_sync(0xffffffffu, bin);
                    if (in && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
                }
                sync();
                if (w == 0) {                                         // lane l owns bins 255-8l .. 248-8l (descending key order)
                    int loc[8], sum = 0;
#pragma unroll
                    for (int b = 0; b < 8; ++b) { loc[b] = hist[255 - 8 * lane - b]; sum += loc[b]; }
                    int incl = sum;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int t2 = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t2; }
                    int accb = incl - sum;
                    if (rank >= accb && rank < incl) {
#pragma unroll
                        for (int b = 0; b < 8; ++b) {
                            if (rank >= accb && rank < accb + loc[b]) { misc[2] = 255 - 8 * lane - b; misc[3] = accb; }
                            accb += loc[b];
                        }
                    }
                }
                sync();
                prefix |= (u64)misc[2] << (8 * byte);
                maskb |= 0xFFull << (8 * byte);
                rank -= misc[3];
            }
        }
        const u64 Pk = want > 0 ? prefix : ~0ull;                      // exactly `want` keys are >= it
        if (tid == 0) misc[0] = 0;
        sync();
        if (!is_final) {
            // survivors (at most 256: the ranks are clamped) go through the shared list, then to the front of the buffer
            for (int i = tid; i < n; i += 256) {
                const u64 key = c[i];
                if (key >= Pk && key != 0ull) { const int pos = atomicAdd(&misc[0], 1); if (pos < 256) list[pos] = key; }
            }
            sync();
            for (int i = tid; i < want && i < 256; i += 256) c[i] = list[i];
            if (tid == 0) {
                if (n_real >= jj) a.tau[q] = Pk;
                a.cnt[q] = want < 256 ? want : 256;
                if (lost) a.status[q] |= AVS_ST_OVERFLOW;
            }
        } else {
            u64* tk = a.topkeys + (size_t)q * a.kprime;
            for (int i = tid; i < n; i += 256) {
                const u64 key = c[i];
                if (key >= Pk && key != 0ull) { const int pos = atomicAdd(&misc[0], 1); if (pos < a.kprime) tk[pos] = key; }
            }
            for (int i = want + tid; i < a.kprime; i += 256) tk[i] = 0ull;
            if (tid == 0) {
                a.cnt[q] = dense_total > 0 ? n : total_in;             // dense level: slots (zero = empty), not keys
                wsel_emit_final(a, q, n_real, Pk, lost);
            }
        }
    }
    sync();
}
