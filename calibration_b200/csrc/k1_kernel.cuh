// Device code of K1 (see k1_fused.cu for the description).  Kept in a header so that tests/host_emul can run
// this very source on the CPU under a lock-step SIMT shim (test-only, CALIB_SIMT_SHIM): there the TMA bulk copy is
// a memcpy, the mbarrier a completion counter and the named barrier a CTA rendezvous; under nvcc those branches
// do not exist.
#pragma once
#include <type_traits>

#include "k1_roles.hpp"
#include "refine_kernels.cuh"
#include "comm_peer.cuh"
#if !defined(CALIB_SIMT_SHIM)
#include "tile_stage.cuh"
#endif

namespace calk {

// ---------------------------------------------------------------------------
// streaming column sums over the 32 lanes of a warp through a padded [32][33] shared tile
// ---------------------------------------------------------------------------
struct LaneSum {
    double* scratch;  // per warp, 32 * 33 doubles
    double* out;      // global: this role's values of this tile
    int lane;
    // values [END - N, END) are complete in the tile: lane j adds the 32 lanes of value END - N + j
    // in a fixed order (four interleaved partial sums, then a fixed tree)
    template <int END, int N>
    __device__ __forceinline__ void flush() {
        __syncwarp();
        if (lane < N) {
            const double* row = scratch + lane * 33;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int k = 0; k < 32; k += 4) { a0 += row[k]; a1 += row[k + 1]; a2 += row[k + 2]; a3 += row[k + 3]; }
            out[END - N + lane] = (a0 + a1) + (a2 + a3);
        }
        __syncwarp();
    }
    template <int K>
    __device__ __forceinline__ void push(double v) {
        scratch[(K & 31) * 33 + lane] = v;
        if constexpr ((K & 31) == 31) flush<K + 1, 32>();
    }
    template <int TOTAL>
    __device__ __forceinline__ void finish() { if constexpr ((TOTAL & 31) != 0) flush<TOTAL, (TOTAL & 31)>(); }
};

struct K1Args {
    DevLayout L;
    EvalBuffers B;
    double huber_delta;
    int nvt;         // values per tile
};

// VIEW: what happens to the products with the view-type pose block (chain rule T_b):
//   0 none (its pose is constant), 1 reduced per tile (bundle: one target pose for all blocks),
//   2 stored per block (intrinsics / extrinsics: per-view unknowns, consumed by the Schur kernels)
enum { VIEW_NONE = 0, VIEW_REDUCE = 1, VIEW_STORE = 2, NOT_FUSED = 3 };

// CTA-shared staging of the tile's observation rows: thread 0 issues bulk asynchronous copies
// (cp.async.bulk, SASS UBLKCP) of RC rows (RC x 1 KB, contiguous in the tile-transposed layout)
// into a two-stage ring; completion is signalled on an mbarrier every thread waits on.  A stage
// is refilled after the CTA barrier that ends its last step (all reads of a step precede it).
template <int RC>
struct CtaStage {
    double* buf;              // [2][RC][4][32]
    unsigned long long* bar;  // [2]
    const double* src;
    int depth, n_chunks;
    static constexpr int kStage = RC * 128;
    static constexpr int kBytes = 2 * kStage * 8 + 16;
    __device__ __forceinline__ void init(unsigned char* smem, const double* tile_src, int tile_depth, bool leader, int n_threads) {
        buf = reinterpret_cast<double*>(smem);
        bar = reinterpret_cast<unsigned long long*>(smem + 2 * kStage * 8);
        src = tile_src; depth = tile_depth; n_chunks = (tile_depth + RC - 1) / RC;
#if defined(CALIB_SIMT_SHIM)   // tests/host_emul: the mbarrier is a completion counter, the named barrier a CTA rendezvous
        if (leader) { simt::mbar_init(&bar[0]); simt::mbar_init(&bar[1]); }
        if (n_threads > 32) simt::named_barrier(n_threads); else __syncwarp();
#else
        if (leader) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        if (n_threads > 32) asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory"); else __syncwarp();
#endif
    }
    __device__ __forceinline__ void issue(int c, bool leader) {
        if (leader && c < n_chunks) {
            const int ks = min(RC, depth - c * RC);
            const unsigned bytes = (unsigned)ks * 1024u;
#if defined(CALIB_SIMT_SHIM)
            simt::bulk_copy_and_complete(buf + (c & 1) * kStage, src + (int64_t)c * kStage, bytes, &bar[c & 1]);
#else
            const unsigned mb = smem_u32(&bar[c & 1]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(buf + (c & 1) * kStage)), "l"(src + (int64_t)c * kStage), "r"(bytes), "r"(mb)
                         : "memory");
#endif
        }
    }
    __device__ __forceinline__ void wait(int c) {
#if defined(CALIB_SIMT_SHIM)
        simt::mbar_wait(&bar[c & 1], (c >> 1) + 1);   // the (c / 2 + 1)-th completion of this stage
#else
        const unsigned mb = smem_u32(&bar[c & 1]);
        const unsigned parity = (unsigned)(c >> 1) & 1u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@!p bra WAIT_%=;\n\t}" ::"r"(mb), "r"(parity) : "memory");
#endif
    }
    __device__ __forceinline__ const double* row(int c, int kk, int lane) const { return buf + (c & 1) * kStage + kk * 128 + lane; }
};

// Values of one observation's Jacobian rows that travel between the roles of a tile: every
// structurally non-zero entry except the constant 1 of the principal-point columns.
template <int MODEL, int IMODE>
struct K1Xchg {
    using LT = Local<MODEL, IMODE>;
    static constexpr bool is_one(int c) { return LT::PI > 0 && (c - 6 == LT::c_cx || c - 6 == LT::c_cy); }
    static constexpr bool send_u(int c) { return LT::has_u(c) && !is_one(c); }
    static constexpr bool send_v(int c) { return LT::has_v(c) && !is_one(c); }
    static constexpr int pos_u(int c) { int n = 0; for (int i = 0; i < c; ++i) n += (send_u(i) ? 1 : 0) + (send_v(i) ? 1 : 0); return n; }
    static constexpr int pos_v(int c) { return pos_u(c) + (send_u(c) ? 1 : 0); }
    static constexpr int NX = pos_u(LT::NL);
};

template <int MODEL, int IMODE, int ROLE>
__device__ __forceinline__ void k1_accumulate(const double* __restrict__ Ju, const double* __restrict__ Jv, double* __restrict__ acc,
                                              double& ssr) {
    using LT = Local<MODEL, IMODE>;
    using RT = K1Roles<MODEL, IMODE>;
    ssr = fma(Ju[LT::NC], Ju[LT::NC], ssr); ssr = fma(Jv[LT::NC], Jv[LT::NC], ssr);
    static_for<0, LT::NL>([&](auto ca) {
        static_for<decltype(ca)::value, LT::NL>([&](auto cb) {
            constexpr int a = decltype(ca)::value, b = decltype(cb)::value;
            constexpr int e = LT::idx(a, b);
            if constexpr (RT::tbl.role[e] == ROLE) {
                constexpr int sl = RT::tbl.slot[e];
                if constexpr (LT::has_u(a) && LT::has_u(b)) acc[sl] = fma(Ju[a], Ju[b], acc[sl]);
                if constexpr (LT::has_v(a) && LT::has_v(b)) acc[sl] = fma(Jv[a], Jv[b], acc[sl]);
            }
        });
    });
}

// Shared memory of one CTA (= one tile, NROLE warps):  [ stage ring | exchange slots ]
template <int MODEL, int IMODE>
struct K1Smem {
    using RT = K1Roles<MODEL, IMODE>;
    static constexpr int NROLE = RT::NROLE;
    static constexpr int RC = NROLE == 3 ? 9 : 8;                 // rows per stage: a multiple of NROLE
    static constexpr int kStageBytes = (CtaStage<RC>::kBytes + 127) / 128 * 128;
    static constexpr int NX = K1Xchg<MODEL, IMODE>::NX;
    static constexpr int kXchgBytes = NROLE > 1 ? 2 * NROLE * NX * 32 * 8 : 0;   // [stage][role][value][lane]
    static constexpr int kScratch = 32 * 33 * 8;                  // epilogue transpose tile, per warp
    static constexpr int kBytes = kStageBytes + (kXchgBytes > NROLE * kScratch ? kXchgBytes : NROLE * kScratch);
};

template <int MODEL, int IMODE, int ROLE, int VIEW>
__device__ __forceinline__ void k1_role(const K1Args& P, int64_t tile, int lane, unsigned char* smem) {
    using LT = Local<MODEL, IMODE>;
    using RT = K1Roles<MODEL, IMODE>;
    using XT = K1Xchg<MODEL, IMODE>;
    using SM = K1Smem<MODEL, IMODE>;
    constexpr int NROLE = RT::NROLE, RC = SM::RC, NX = XT::NX;
    constexpr int NA = RT::count(ROLE);
    constexpr int NL = LT::NL, NC = LT::NC, PI = LT::PI;
    const DevLayout& L = P.L; const EvalBuffers& B = P.B;
    const int64_t s = tile * 32 + lane;
    const int len = L.seg_len[s];
    const int depth = L.tile_depth[tile];
    const bool leader = ROLE == 0 && lane == 0;
    CtaStage<RC> ts; ts.init(smem, L.obs + L.tile_off[tile] * 128, depth, leader, NROLE * 32);
    ts.issue(0, leader); ts.issue(1, leader);
    double* xbuf = reinterpret_cast<double*>(smem + SM::kStageBytes);   // [2][NROLE][NX][32]
    double A[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) A[i] = B.seg_frame[(int64_t)i * L.n_seg + s];
    const CamConst c = B.camc[L.seg_cam[s]];
    double acc[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) acc[i] = 0.0;
    double ssr = 0.0;
    // UNI: every lane of this warp has a full-depth segment and the depth is a multiple of NROLE (the
    // benchmark's boards: 88 corners everywhere) — the step loop then carries no per-lane predicates at all.
    // The barrier count per step is the same on both paths, so the roles of a tile may choose differently.
    auto run = [&](auto uni_c) {
        constexpr bool UNI = decltype(uni_c)::value;
        int step = 0;
        for (int ch = 0; ch < ts.n_chunks; ++ch) {
            ts.wait(ch);
            const int k0 = ch * RC, kn = min(RC, depth - k0);
            for (int kk = 0; kk < kn; kk += NROLE, ++step) {
                // each role projects one of the NROLE corners of this step ...
                const int kmine = kk + ROLE;
                const bool mine = UNI || k0 + kmine < len;
                double Ju[NL], Jv[NL];
                if (mine) {
                    const double* q = ts.row(ch, kmine, lane);
                    obs_rows<MODEL, IMODE>(c, A, q[0], q[32], q[64], q[96], Ju, Jv);
                }
                if constexpr (NROLE > 1) {
                    // ... hands the rows to the other roles through shared memory ...
                    double* xs = xbuf + ((step & 1) * NROLE + ROLE) * NX * 32 + lane;
                    if (mine) {
                        static_for<0, NL>([&](auto cc) {
                            constexpr int col = decltype(cc)::value;
                            if constexpr (XT::send_u(col)) xs[XT::pos_u(col) * 32] = Ju[col];
                            if constexpr (XT::send_v(col)) xs[XT::pos_v(col) * 32] = Jv[col];
                        });
                    }
                }
                // ... accumulates ITS entries of the local system for its own corner (the published rows
                // land in shared memory meanwhile, and the partner has time to arrive at the barrier) ...
                if (mine) k1_accumulate<MODEL, IMODE, ROLE>(Ju, Jv, acc, ssr);
                if constexpr (NROLE > 1) {
#if defined(CALIB_SIMT_SHIM)
                    simt::named_barrier(NROLE * 32);
#else
                    asm volatile("bar.sync 1, %0;" ::"n"(NROLE * 32) : "memory");
#endif
                    // ... and for the corners the other roles projected
                    static_for<1, NROLE>([&](auto cp) {
                        constexpr int other = (ROLE + decltype(cp)::value) % NROLE;
                        if (UNI || k0 + kk + other < len) {
                            const double* xo = xbuf + ((step & 1) * NROLE + other) * NX * 32 + lane;
                            double Pu[NL], Pv[NL];
                            static_for<0, NL>([&](auto cc) {
                                constexpr int col = decltype(cc)::value;
                                Pu[col] = XT::send_u(col) ? xo[XT::pos_u(col) * 32] : (LT::has_u(col) ? 1.0 : 0.0);
                                Pv[col] = XT::send_v(col) ? xo[XT::pos_v(col) * 32] : (LT::has_v(col) ? 1.0 : 0.0);
                            });
                            k1_accumulate<MODEL, IMODE, ROLE>(Pu, Pv, acc, ssr);
                        }
                    });
                }
            }
            if constexpr (NROLE == 1) __syncwarp();  // every lane is done with this stage (NROLE > 1: the step barrier)
            ts.issue(ch + 2, leader);                // refill it
        }
    };
    if (__all_sync(0xffffffffu, len == depth) && depth % NROLE == 0) run(std::true_type{}); else run(std::false_type{});
    unsigned char* warp_smem = smem;             // NROLE == 1: the transpose tile aliases the idle staging ring
    if constexpr (NROLE > 1) {
        // the exchange slots are dead: they become the transpose tiles
#if defined(CALIB_SIMT_SHIM)
        simt::named_barrier(NROLE * 32);
#else
        asm volatile("bar.sync 1, %0;" ::"n"(NROLE * 32) : "memory");
#endif
        warp_smem = smem + SM::kStageBytes + ROLE * SM::kScratch;
    }
    if constexpr (VIEW == NOT_FUSED) {
        static_for<0, LT::NE>([&](auto ce) {
            constexpr int e = decltype(ce)::value;
            if constexpr (RT::tbl.role[e] == ROLE) { constexpr int sl = RT::tbl.slot[e]; B.segN[(int64_t)e * L.n_seg + s] = acc[sl]; }
        });
        if (ROLE == 0) B.segN[(int64_t)RT::RR * L.n_seg + s] = ssr;
    } else {
        // ---------------- fused epilogue: lane = residual block (s == device block id) ----------------
        constexpr bool RED = VIEW == VIEW_REDUCE;
        constexpr int K_RR = NA;                                    // role 0: rr, cost
        constexpr int K_EVI = NA + (ROLE == 0 ? 2 : 0);             // owned E_vi columns, 6 values each
        constexpr int K_GV = K_EVI + 6 * RT::n_owned_cols(ROLE);    // role 0: g_v(6) H_vv(21) Q(36)
        constexpr int TOTAL = RT::n_vals(ROLE, RED);
        double rho, w; huber_weight(P.huber_delta, ssr, rho, w);
        LaneSum ls{reinterpret_cast<double*>(warp_smem), B.tile_vals + tile * P.nvt + RT::val_off(ROLE, RED), lane};
        static_for<0, NA>([&](auto ci) { constexpr int i = decltype(ci)::value; ls.template push<i>(w * acc[i]); });
        if constexpr (ROLE == 0) {
            ls.template push<K_RR>(w * ssr); ls.template push<K_RR + 1>(0.5 * rho);
            B.blk_ssr[s] = ssr;
        }
        if constexpr (VIEW != VIEW_NONE && (ROLE == 0 || RT::owns_cols(ROLE))) {
            const int64_t nb = L.n_blk;
            const bool vfree = RED ? true : (L.blk_vfree[s] != 0);
            // T = [[TL, 0], [BL, TL / 2]] (view_transform, k1_math.cuh): 18 loads, structural zeros known to the compiler
            double T[36];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int cix = 0; cix < 3; ++cix) T[6 * r + cix] = B.blk_Tv[(int64_t)(6 * r + cix) * nb + s];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cix = 0; cix < 3; ++cix) { T[6 * r + 3 + cix] = 0.0; T[6 * (r + 3) + 3 + cix] = 0.5 * T[6 * r + cix]; }
            // E_vi columns of this role: w T^T N_xi,i[:, j]
            static_for<0, PI>([&](auto cj) {
                constexpr int j = decltype(cj)::value;
                if constexpr (RT::tbl.col_role[j] == ROLE) {
                    static_for<0, 6>([&](auto ci) {
                        constexpr int i = decltype(ci)::value;
                        double a = 0.0;
                        static_for<0, 6>([&](auto ck) {
                            constexpr int k = decltype(ck)::value;
                            constexpr int sl = RT::tbl.slot[LT::idx(k, 6 + j)];
                            a = fma(T[6 * k + i], acc[sl], a);
                        });
                        a *= w;
                        if constexpr (RED) ls.template push<K_EVI + 6 * RT::col_rank(j) + i>(a);
                        else if (vfree) B.blk_Evi[(int64_t)(PI * i + j) * nb + s] = a;
                    });
                }
            });
            if constexpr (ROLE == 0) {
                static_for<0, 6>([&](auto ci) {
                    constexpr int i = decltype(ci)::value;
                    double a = 0.0;
                    static_for<0, 6>([&](auto ck) {
                        constexpr int k = decltype(ck)::value;
                        constexpr int sl = RT::tbl.slot[LT::idx(k, NC)];
                        a = fma(T[6 * k + i], acc[sl], a);
                    });
                    a *= w;
                    if constexpr (RED) ls.template push<K_GV + i>(a);
                    else if (vfree) B.blk_gv[(int64_t)i * nb + s] = a;
                });
                double Q[36];  // w T^T N_xixi
                static_for<0, 6>([&](auto ci) {
                    static_for<0, 6>([&](auto cjj) {
                        constexpr int i = decltype(ci)::value, j = decltype(cjj)::value;
                        double a = 0.0;
                        static_for<0, 6>([&](auto ck) {
                            constexpr int k = decltype(ck)::value;
                            constexpr int sl = RT::tbl.slot[k <= j ? LT::idx(k, j) : LT::idx(j, k)];
                            a = fma(T[6 * k + i], acc[sl], a);
                        });
                        Q[6 * i + j] = a * w;
                    });
                });
                static_for<0, 6>([&](auto ci) {
                    static_for<decltype(ci)::value, 6>([&](auto cjj) {
                        constexpr int i = decltype(ci)::value, j = decltype(cjj)::value;
                        constexpr int o = i * 6 - i * (i - 1) / 2 + (j - i);
                        double a = 0.0;
#pragma unroll
                        for (int k = 0; k < 6; ++k) a = fma(Q[6 * i + k], T[6 * k + j], a);
                        if constexpr (RED) ls.template push<K_GV + 6 + o>(a);
                        else if (vfree) B.blk_Hvv[(int64_t)o * nb + s] = a;
                    });
                });
                if constexpr (RED) {
                    static_for<0, 36>([&](auto ci) { constexpr int i = decltype(ci)::value; ls.template push<K_GV + 27 + i>(Q[i]); });
                } else {
                    const double* __restrict__ Tc = B.camT + (int64_t)L.seg_cam[s] * 36;
#pragma unroll
                    for (int i = 0; i < 6; ++i)
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            double a = 0.0;
#pragma unroll
                            for (int k = 0; k < 6; ++k) a = fma(Q[6 * i + k], Tc[6 * k + j], a);
                            if (vfree) B.blk_Evc[(int64_t)(6 * i + j) * nb + s] = a;
                        }
                }
            }
        }
        ls.template finish<TOTAL>();
    }
}

template <int MODEL, int IMODE, int VIEW>
__global__ void __launch_bounds__(K1Roles<MODEL, IMODE>::NROLE * 32) k1_kernel(const __grid_constant__ K1Args P) {
    using RT = K1Roles<MODEL, IMODE>;
    // one tile per CTA, one warp per role: small CTAs drift apart in time, so the epilogue of one
    // overlaps the main loops of its neighbours on the SM
    const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = blockIdx.x;
    extern __shared__ __align__(128) unsigned char k1_smem[];
    if (role == 0) k1_role<MODEL, IMODE, 0, VIEW>(P, tile, lane, k1_smem);
    if constexpr (RT::NROLE > 1) { if (role == 1) k1_role<MODEL, IMODE, 1, VIEW>(P, tile, lane, k1_smem); }
    if constexpr (RT::NROLE > 2) { if (role == 2) k1_role<MODEL, IMODE, 2, VIEW>(P, tile, lane, k1_smem); }
}

// ---------------------------------------------------------------------------
// per-camera sums of the per-tile rows, fixed order, no floating-point atomics; launcher: k1_fused.cu.
// ONE launch: every CTA sums the rows of its chunk (32 tiles of one camera); the CTA that finishes a camera's last
// chunk (an integer ticket per camera) adds that camera's chunk partials in chunk order; the CTA that finishes the
// last camera resets the tickets and, on several GPUs, runs the NVLink peer-memory all-reduce of cam_sums itself
// (comm_peer.cuh) — the reduction and the collective that follows it are one kernel.  Which CTA happens to be last
// does not matter: every sum is formed by one thread in a fixed order, so results are run-to-run identical.
// ---------------------------------------------------------------------------
struct TileReduceArgs {
    const double* tile_vals; int nvt;
    const ColChunk* chunks; double* partial;          // [n_chunks][nvt]
    const int32_t* cam_chunk_off; int n_cams;         // [n_cams + 1]
    const int32_t* vmap; double* cam_sums; int NV;    // value v of a row -> cam_sums[cam * NV + vmap[v]]
    unsigned* tickets;                                // [n_cams + 1], zero between launches
    int n_active_cams;                                // cameras that own at least one chunk on this GPU
};
__global__ void __launch_bounds__(256) k_tile_reduce(const TileReduceArgs A, const calcomm::PeerArgs peer) {
#if defined(__CUDACC__)
    __shared__ int s_last, s_bad;
#else
    static int s_last, s_bad;   // host build of this source (tests/host_emul): one CTA runs at a time
#endif
    const ColChunk c = A.chunks[blockIdx.x];
    const int nvt = A.nvt;
    for (int v = threadIdx.x; v < nvt; v += 256) {
        double a = 0.0;
        for (int64_t t = c.begin; t < c.end; ++t) a += A.tile_vals[t * nvt + v];
        A.partial[(int64_t)blockIdx.x * nvt + v] = a;
    }
    __threadfence();
    __syncthreads();
    const int cam = c.cam, c0 = A.cam_chunk_off[cam], c1 = A.cam_chunk_off[cam + 1];
    if (threadIdx.x == 0) s_last = atomicAdd(&A.tickets[cam], 1u) == (unsigned)(c1 - c0 - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int v = threadIdx.x; v < nvt; v += 256) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;   // four interleaved partial sums over the chunks, fixed order
        int k = c0;
        for (; k + 3 < c1; k += 4) {
            a0 += __ldcg(A.partial + (int64_t)k * nvt + v); a1 += __ldcg(A.partial + (int64_t)(k + 1) * nvt + v);
            a2 += __ldcg(A.partial + (int64_t)(k + 2) * nvt + v); a3 += __ldcg(A.partial + (int64_t)(k + 3) * nvt + v);
        }
        for (; k < c1; ++k) a0 += __ldcg(A.partial + (int64_t)k * nvt + v);
        A.cam_sums[(int64_t)cam * A.NV + A.vmap[v]] = (a0 + a1) + (a2 + a3);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&A.tickets[A.n_cams], 1u) == (unsigned)(A.n_active_cams - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int i = threadIdx.x; i <= A.n_cams; i += 256) A.tickets[i] = 0u;   // ready for the next launch
#if !defined(CALIB_SIMT_SHIM)
    if (peer.world > 1) calcomm::peer_allreduce_cta(A.cam_sums, A.n_cams * A.NV, peer, &s_bad);
#endif
    (void)s_bad;
}

}  // namespace calk
