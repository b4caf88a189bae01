// Host emulation of the filter / smoother kernels (TEST INFRASTRUCTURE, CPU only).
//
// The per-chunk bodies in csrc/filter_impl.cuh and csrc/smoother_impl.cuh are __host__ __device__;
// this harness runs them in plain host loops (one iteration per GPU thread) with the same chunk
// planning, SoA workspace layout, carry fold and first/last-step rules as the CUDA drivers, so the
// arithmetic and indexing of the product kernels can be checked against the oracle on a machine
// with no GPU.  Only the warp-shuffle scan kernel is replaced (by a sequential fold with the same
// Alg::combine).  Never loaded by the product.
#include <vector>
#include <cstring>
#include "../../bayesnewton_b200/csrc/filter_impl.cuh"
#include "../../bayesnewton_b200/csrc/smoother_impl.cuh"
#include "../../bayesnewton_b200/csrc/sites_impl.cuh"
#include "../../bayesnewton_b200/csrc/up_impl.cuh"
#include "../../bayesnewton_b200/csrc/iter_impl.cuh"
#include "../../bayesnewton_b200/csrc/gd_impl.cuh"

namespace bn {
void set_error(const char*, ...) {}

template <class Alg>
static void host_scan(const double* in, long long n, double* prefix) {
    typename Alg::Elem acc, e, r;
    for (long long i = 0; i < n; ++i) {
        Alg::load(in, n, i, e);
        if (i == 0) acc = e; else { Alg::combine(acc, e, r); acc = r; }
        Alg::store(prefix, n, i, acc);
    }
}

// the scan as the fused iteration does it (scan.cuh: warp_prescan + scan_last_levels_kernel): inclusive prefixes inside
// groups of 32 consecutive elements, inclusive prefixes of the warp totals inside groups of kEmuGroup warps, and inclusive
// prefixes of the group totals (the device groups 288 warps; a small group here reaches all three parts at test sizes)
constexpr int kEmuGroup = 3;
struct HostPrefix {
    std::vector<double> warp, group;
    long long nw = 0, ng = 0;
    PrefixParts parts() const {
        PrefixParts u;
        if (nw > 1) { u.warp = warp.data(); u.n_warp = nw; }
        if (ng > 1) { u.group = group.data(); u.n_group = ng; u.gsize = kEmuGroup; }
        return u;
    }
    // (array, count) holding the grand total as its last element
    template <class F> auto total(const double* prefix0, long long n, F f) const {
        if (ng > 1) return f(group.data(), ng);
        if (nw > 1) return f(warp.data(), nw);
        return f(prefix0, n);
    }
};
template <class Alg>
static void host_warp_scan(const double* in, long long n, double* prefix0, HostPrefix& hp) {
    hp.nw = (n + 31) / 32;
    hp.ng = (hp.nw + kEmuGroup - 1) / kEmuGroup;
    hp.warp.assign((size_t)hp.nw * Alg::kElem, 0.0);
    hp.group.assign((size_t)hp.ng * Alg::kElem, 0.0);
    typename Alg::Elem acc, e, r, wacc, gacc;
    for (long long i = 0; i < n; ++i) {
        Alg::load(in, n, i, e);
        if ((i & 31) == 0) acc = e; else { Alg::combine(acc, e, r); acc = r; }
        Alg::store(prefix0, n, i, acc);
        if ((i & 31) == 31 || i == n - 1) {
            const long long w = i >> 5;
            if (w % kEmuGroup == 0) wacc = acc; else { Alg::combine(wacc, acc, r); wacc = r; }
            Alg::store(hp.warp.data(), hp.nw, w, wacc);
            if (w % kEmuGroup == kEmuGroup - 1 || i == n - 1) {
                const long long g = w / kEmuGroup;
                if (g == 0) gacc = wacc; else { Alg::combine(gacc, wacc, r); gacc = r; }
                Alg::store(hp.group.data(), hp.ng, g, gacc);
            }
        }
    }
    if (hp.ng <= 1) {  // a single group: its within-group prefixes are the global ones (the device's two-level plans)
        hp.ng = 0;
    }
}

// time-sharded filter over `world` emulated ranks
template <class MakeGen>
static int emu_kf(MakeGen make, int d_unused, int form, long long N, int L, int world, const double* y, int D,
                  const double* R, const unsigned char* mask, int return_predict, double* ell, double* fms,
                  double* fPs, int dstate) {
    (void)d_unused;
    using Gen = decltype(make(0LL));
    constexpr int d = Gen::d;
    using Alg = FilterAlg<d>;
    if (form == BN_SEQUENTIAL) {
        Gen gen = make(0LL);
        KfIO io{N, y, R, mask, fms, fPs, return_predict};
        if (ell) kf_seq_body<Gen, true>(gen, io, ell); else kf_seq_body<Gen, false>(gen, io, nullptr);
        return 0;
    }
    std::vector<long long> off(world + 1);
    for (int r = 0; r <= world; ++r) off[r] = N * r / world;
    std::vector<double> carries((size_t)world * Alg::kCarry);
    std::vector<std::vector<double>> prefixes(world);
    std::vector<long long> nch(world);
    auto shard_io = [&](int r) {
        long long o = off[r];
        return KfIO{off[r + 1] - o, y + o * D, R + o * D * D, mask ? mask + o * D : nullptr,
                    fms ? fms + o * dstate : nullptr, fPs ? fPs + o * dstate * dstate : nullptr, return_predict};
    };
    for (int r = 0; r < world; ++r) {
        Gen gen = make(off[r]);
        KfIO io = shard_io(r);
        long long nc = (io.N + L - 1) / L;
        nch[r] = nc;
        std::vector<double> agg((size_t)nc * Alg::kElem);
        prefixes[r].assign((size_t)nc * Alg::kElem, 0.0);
        for (long long c = 0; c < nc; ++c) kf_reduce_chunk(gen, io, L, nc, r == 0, agg.data(), c);
        host_scan<Alg>(agg.data(), nc, prefixes[r].data());
        export_carry_body<Alg>(prefixes[r].data(), nc, carries.data() + (size_t)r * Alg::kCarry);
    }
    double total = 0.0;
    for (int r = 0; r < world; ++r) {
        Gen gen = make(off[r]);
        KfIO io = shard_io(r);
        double s0[64];
        fold_carries_body<Alg>(carries.data(), 0, r, 1, s0);
        std::vector<double> partials(nch[r]);
        for (long long c = 0; c < nch[r]; ++c) {
            if (ell) kf_apply_chunk<Gen, true>(gen, io, L, nch[r], r == 0, prefixes[r].data(), s0, partials.data(), c);
            else kf_apply_chunk<Gen, false>(gen, io, L, nch[r], r == 0, prefixes[r].data(), s0, nullptr, c);
        }
        for (long long c = 0; c < nch[r]; ++c) total += partials[c];
    }
    if (ell) *ell = total;
    return 0;
}

template <class MakeGen>
static int emu_rts(MakeGen make, int form, long long N, int L, int world, const double* fms, const double* fPs,
                   int return_full, double* sms, double* sPs, double* gains) {
    using Gen = decltype(make(0LL));
    constexpr int d = Gen::d, Df = Gen::D;
    using Alg = SmootherAlg<d>;
    const int od = return_full ? d : Df;
    if (form == BN_SEQUENTIAL) {
        Gen gen = make(0LL);
        RtsIO io{N, fms, fPs, sms, sPs, gains, return_full};
        rts_seq_body(gen, io);
        return 0;
    }
    std::vector<long long> off(world + 1);
    for (int r = 0; r <= world; ++r) off[r] = N * r / world;
    std::vector<double> carries((size_t)world * Alg::kCarry);
    std::vector<std::vector<double>> prefixes(world);
    std::vector<long long> nch(world);
    auto shard_io = [&](int r) {
        long long o = off[r];
        return RtsIO{off[r + 1] - o, fms + o * d, fPs + o * d * d, sms + o * od, sPs + o * od * od,
                     gains ? gains + o * d * d : nullptr, return_full};
    };
    for (int r = 0; r < world; ++r) {
        Gen gen = make(off[r]);
        RtsIO io = shard_io(r);
        long long nc = (io.N + L - 1) / L;
        nch[r] = nc;
        std::vector<double> agg((size_t)nc * Alg::kElem);
        prefixes[r].assign((size_t)nc * Alg::kElem, 0.0);
        for (long long c = 0; c < nc; ++c) rts_reduce_chunk(gen, io, L, nc, r == world - 1, agg.data(), c);
        host_scan<Alg>(agg.data(), nc, prefixes[r].data());
        export_carry_body<Alg>(prefixes[r].data(), nc, carries.data() + (size_t)r * Alg::kCarry);
    }
    for (int r = 0; r < world; ++r) {
        Gen gen = make(off[r]);
        RtsIO io = shard_io(r);
        double s0[64];
        fold_carries_body<Alg>(carries.data(), world - 1, r, -1, s0);
        for (long long c = 0; c < nch[r]; ++c)
            rts_apply_chunk(gen, io, L, nch[r], r == world - 1, prefixes[r].data(), s0, c);
    }
    return 0;
}


// fused posterior update (up_impl.cuh) over `world` emulated ranks: the three sharded phases with
// DirectCtx IO, host scans, the same tiled scratch layout, carry export and folds as up_run().
template <class G>
static int emu_up(const bn_kernel_spec* k, long long N, int L, int world, const double* dt, const double* y,
                  const double* R, const unsigned char* mask, double* ell, double* pm, double* pc,
                  double* dvar = nullptr, double* dlen = nullptr) {
    constexpr int d = G::d, D = G::D;
    using FA = FilterAlg<d>;
    using SA = SmootherAlg<d>;
    if (L % kUpTJ != 0) return -2;
    G g;
    g.prepare(*k);
    std::vector<long long> off(world + 1);
    for (int r = 0; r <= world; ++r) off[r] = N * r / world;
    struct Rank { long long n, nc; std::vector<double> agg, fpre, sel, spre, fs, s0, sinit; UpIO io; };
    std::vector<Rank> ranks(world);
    std::vector<double> fcar((size_t)world * FA::kCarry), scar((size_t)world * SA::kCarry);
    for (int r = 0; r < world; ++r) {
        Rank& q = ranks[r];
        long long o = off[r];
        q.n = off[r + 1] - o;
        q.nc = (q.n + L - 1) / L;
        q.io = UpIO{q.n, dt + o, y + o * D, R + o * D * D, mask ? mask + o * D : nullptr, pm + o * D, pc + o * D * D};
        q.agg.assign((size_t)q.nc * FA::kElem, 0.0);
        q.fpre.assign((size_t)q.nc * FA::kElem, 0.0);
        q.sel.assign((size_t)q.nc * SA::kElem, 0.0);
        q.spre.assign((size_t)q.nc * SA::kElem, 0.0);
        q.fs.assign((size_t)fs_doubles(q.nc, L, d + symn(d)), 0.0);
        q.s0.assign(64, 0.0);
        q.sinit.assign(64, 0.0);
        DirectCtx<D> cx{q.io};
        for (long long c = 0; c < q.nc; ++c) up_reduce_chunk(g, cx, q.n, L, q.nc, r == 0, q.agg.data(), c, true);
        host_scan<FA>(q.agg.data(), q.nc, q.fpre.data());
        export_carry_body<FA>(q.fpre.data(), q.nc, fcar.data() + (size_t)r * FA::kCarry);
    }
    double total = 0.0;
    for (int r = 0; r < world; ++r) {
        Rank& q = ranks[r];
        DirectCtx<D> cx{q.io};
        fold_carries_body<FA>(fcar.data(), 0, r, 1, q.s0.data());
        std::vector<double> partials(q.nc, 0.0);
        for (long long c = 0; c < q.nc; ++c) {
            if (ell) up_filter_chunk<G, true>(g, cx, q.n, L, q.nc, r == 0, q.fpre.data(), q.s0.data(), q.fs.data(),
                                              partials.data(), c, true);
            else up_filter_chunk<G, false>(g, cx, q.n, L, q.nc, r == 0, q.fpre.data(), q.s0.data(), q.fs.data(),
                                           nullptr, c, true);
        }
        for (double v : partials) total += v;
        for (long long c = 0; c < q.nc; ++c)
            up_selem_chunk<G>(q.n, L, q.nc, r != 0, q.agg.data(), q.s0.data(), q.fs.data(), q.sel.data(), c);
        host_scan<SA>(q.sel.data(), q.nc, q.spre.data());
        up_export_scarry<d>(q.spre.data(), q.nc, r == world - 1, q.n, L, q.fs.data(),
                            scar.data() + (size_t)r * SA::kCarry);
    }
    if (ell) *ell = total;
    if (dvar) for (int c = 0; c < G::NC; ++c) dvar[c] = dlen[c] = 0.0;
    for (int r = 0; r < world; ++r) {
        Rank& q = ranks[r];
        DirectCtx<D> cx{q.io};
        if (r != world - 1) fold_carries_body<SA>(scar.data(), world - 1, r, -1, q.sinit.data());
        else up_last_state<d>(q.n, L, q.fs.data(), q.sinit.data());
        if (dvar) {  // hyper-gradient: per-chunk fields -> per-rank sums -> chain rule; ranks add up (all-reduce)
            constexpr int NF = GradAcc<G>::kFields;
            std::vector<double> gpart((size_t)q.nc * NF, 0.0), fields(NF, 0.0);
            for (long long c = 0; c < q.nc; ++c)
                up_smooth_chunk<G, true>(g, cx, q.n, L, q.nc, q.spre.data(), q.sinit.data(), q.fs.data(), c, true,
                                         r == 0, q.s0.data(), gpart.data());
            for (int f = 0; f < NF; ++f)
                for (long long c = 0; c < q.nc; ++c) fields[f] += gpart[(size_t)f * q.nc + c];
            double dv[G::NC], dl[G::NC];
            grad_finish<G>(*k, fields.data(), dv, dl);
            for (int c = 0; c < G::NC; ++c) { dvar[c] += dv[c]; dlen[c] += dl[c]; }
        } else {
            for (long long c = 0; c < q.nc; ++c)
                up_smooth_chunk<G, false>(g, cx, q.n, L, q.nc, q.spre.data(), q.sinit.data(), q.fs.data(), c, true);
        }
    }
    return 0;
}


// ---- fused iteration pass on tiled state (iter_impl.cuh) over `world` emulated ranks.  The caller's arrays are
// linear; each rank tiles its shard (the role of bn_iter_to_tiled), runs the three phases with the tiled chunk bodies
// and the requested smoother epilogue, and the results are un-tiled (bn_iter_from_tiled).
template <typename T>
static void host_to_tiled(long long n, int L, long long nc, const T* x, std::vector<T>& xt, T fill) {
    xt.assign((size_t)tl_len(nc, L), fill);
    for (long long c = 0; c < nc; ++c)
        for (int j = 0; j < L; ++j) {
            const long long k = c * L + j;
            if (k < n) xt[(size_t)(tl_base(c, L) + 32LL * j)] = x[k];
        }
}
static void host_from_tiled(long long n, int L, long long nc, const std::vector<double>& xt, double* x) {
    for (long long c = 0; c < nc; ++c)
        for (int j = 0; j < L; ++j) {
            const long long k = c * L + j;
            if (k < n) x[k] = xt[(size_t)(tl_base(c, L) + 32LL * j)];
        }
}

template <class G, template <int, int, bool> class Epi, int LIK, int METHOD, bool TAB>
static void emu_it_sweep(const G& g, const ItIO& io, const ItSiteArgs& sa, const Cub1& cub, int L, long long nc,
                         const double* spre, const double* sinit, const double* fs, const PrefixParts& swp) {
    const double* tab = TAB ? probit_table_host().data() : nullptr;
    for (long long c = 0; c < nc; ++c) {
        Epi<LIK, METHOD, TAB> epi(io, sa, &cub, tab);
        it_smooth_chunk(g, io, L, nc, spre, sinit, fs, c, epi, swp);
    }
}

template <class G>
static int emu_it(const bn_kernel_spec* k, long long N, int L, int world, const double* dt, const double* y,
                  double* sy, double* sR, const unsigned char* mask, int mode, int method, int lik, double lik_param,
                  int Q, const double* cx, const double* cw, double lr, double power, int ensure_psd, int use_table,
                  double* ell, double* sums, double* pm, double* pc, int spec = 0, double* jstar_mean = nullptr) {
    constexpr int d = G::d;
    using FA = FilterAlg<d>;
    using SA = SmootherAlg<d>;
    if (L % kUpTJ != 0) return -2;
    G g;
    g.prepare(*k);
    Cub1 cub;
    make_cub1(Q, cx, cw, cub);
    std::vector<long long> off(world + 1);
    for (int r = 0; r <= world; ++r) off[r] = N * r / world;
    struct Rank {
        long long n, nc;
        std::vector<double> agg, fpre, sel, spre, fs, s0, sinit, dt_t, y_t, sy_t, sR_t, pm_t, pc_t, p1, p2, ell1;
        HostPrefix fhp, shp;
        PrefixParts fw() const { return fhp.parts(); }
        PrefixParts sw() const { return shp.parts(); }
        std::vector<unsigned char> mk_t;
        std::vector<int> jst;
        ItIO io;
    };
    std::vector<Rank> ranks(world);
    std::vector<double> fcar((size_t)world * FA::kCarry), scar((size_t)world * SA::kCarry);
    for (int r = 0; r < world; ++r) {
        Rank& q = ranks[r];
        const long long o = off[r];
        q.n = off[r + 1] - o;
        q.nc = (q.n + L - 1) / L;
        host_to_tiled<double>(q.n, L, q.nc, dt + o, q.dt_t, 0.0);
        host_to_tiled<double>(q.n, L, q.nc, y + o, q.y_t, 0.0);
        host_to_tiled<double>(q.n, L, q.nc, sy + o, q.sy_t, 0.0);
        host_to_tiled<double>(q.n, L, q.nc, sR + o, q.sR_t, 1.0);
        if (mask) host_to_tiled<unsigned char>(q.n, L, q.nc, mask + o, q.mk_t, (unsigned char)0);
        q.pm_t.assign((size_t)tl_len(q.nc, L), 0.0);
        q.pc_t.assign((size_t)tl_len(q.nc, L), 0.0);
        q.io = ItIO{q.n, q.dt_t.data(), q.y_t.data(), q.sy_t.data(), q.sR_t.data(), mask ? q.mk_t.data() : nullptr,
                    q.pm_t.data(), q.pc_t.data(), nullptr, nullptr};
        q.agg.assign((size_t)q.nc * FA::kElem, 0.0);
        q.fpre.assign((size_t)q.nc * FA::kElem, 0.0);
        q.sel.assign((size_t)q.nc * SA::kElem, 0.0);
        q.spre.assign((size_t)q.nc * SA::kElem, 0.0);
        q.fs.assign((size_t)fs_doubles(q.nc, L, d + symn(d)) + 32 * (d + symn(d)), 0.0);
        q.s0.assign(64, 0.0);
        q.sinit.assign(64, 0.0);
        q.p1.assign(q.nc, 0.0);
        q.p2.assign(q.nc, 0.0);
        q.jst.assign(q.nc, 0);
        q.ell1.assign(q.nc, 0.0);
        if (spec) {  // SpecReduce: the 32 chunks of a warp advance in lockstep and switch together
            for (long long c0 = 0; c0 < q.nc; c0 += 32) {
                std::vector<SpecReduce<G, true>> sr(32);
                for (int l = 0; l < 32; ++l) sr[l].init(g, q.io, L, q.fs.data(), c0 + l < q.nc ? c0 + l : 0, c0 + l < q.nc);
                int j = 0;
                bool all_dec = false;
                while (!all_dec) {
                    all_dec = true;
                    for (int l = 0; l < 32; ++l) {
                        bool dec = true;
                        if (j < sr[l].cnt) dec = sr[l].absorb(g, r == 0 && c0 + l == 0 && j == 0);
                        all_dec = all_dec && dec;
                    }
                    ++j;
                }
                for (int l = 0; l < 32 && c0 + l < q.nc; ++l) {
                    q.jst[c0 + l] = j < sr[l].cnt ? j : sr[l].cnt;
                    FA::store(q.agg.data(), q.nc, c0 + l, sr[l].el);
                    for (int jj = j; jj < sr[l].cnt; ++jj) sr[l].filter(g);
                    if (q.jst[c0 + l] < sr[l].cnt) sr[l].store_state(q.agg.data(), q.nc, c0 + l);
                    q.ell1[c0 + l] = sr[l].ell;
                }
            }
        } else {
            for (long long c = 0; c < q.nc; ++c) it_reduce_chunk(g, q.io, L, q.nc, r == 0, q.agg.data(), c);
        }
        host_warp_scan<FA>(q.agg.data(), q.nc, q.fpre.data(), q.fhp);
        q.fhp.total(q.fpre.data(), q.nc, [&](const double* a, long long n) {
            export_carry_body<FA>(a, n, fcar.data() + (size_t)r * FA::kCarry);
            return 0;
        });
    }
    double total = 0.0;
    for (int r = 0; r < world; ++r) {
        Rank& q = ranks[r];
        fold_carries_body<FA>(fcar.data(), 0, r, 1, q.s0.data());
        std::vector<double> partials(q.nc, 0.0);
        for (long long c = 0; c < q.nc; ++c)
            it_filter_chunk<G, true>(g, q.io, L, q.nc, r == 0, q.fpre.data(), q.s0.data(), q.fs.data(), partials.data(), c,
                                     spec ? q.jst.data() : nullptr, q.fw());
        for (double v : partials) total += v;
        for (double v : q.ell1) total += v;
        if (jstar_mean) { double a = 0; for (int v : q.jst) a += v; *jstar_mean += a / (double)q.nc / world; }
        for (long long c = 0; c < q.nc; ++c)
            up_selem_chunk<G>(q.n, L, q.nc, r != 0, q.agg.data(), q.s0.data(), q.fs.data(), q.sel.data(), c);
        host_warp_scan<SA>(q.sel.data(), q.nc, q.spre.data(), q.shp);
        q.shp.total(q.spre.data(), q.nc, [&](const double* a, long long n) {
            up_export_scarry<d>(a, n, r == world - 1, q.n, L, q.fs.data(), scar.data() + (size_t)r * SA::kCarry);
            return 0;
        });
    }
    if (ell) *ell = total;
    double s1 = 0.0, s2 = 0.0;
    for (int r = 0; r < world; ++r) {
        Rank& q = ranks[r];
        if (r != world - 1) fold_carries_body<SA>(scar.data(), world - 1, r, -1, q.sinit.data());
        else up_last_state<d>(q.n, L, q.fs.data(), q.sinit.data());
        ItSiteArgs sa{lik_param, lr, power, ensure_psd, 0, q.p1.data(), q.p2.data()};
        bool done = false;
        if (mode == IT_PLAIN) {
            for (long long c = 0; c < q.nc; ++c) {
                EpiStore epi{q.io.pm, q.io.pc, nullptr, nullptr};
                it_smooth_chunk(g, q.io, L, q.nc, q.spre.data(), q.sinit.data(), q.fs.data(), c, epi, q.sw());
            }
            done = true;
        }
#define X(LK, M)                                                                                                          \
        if (!done && lik == LK && method == M) {                                                                          \
            constexpr bool kTab = (LK == BN_LIK_BERNOULLI_PROBIT && M == BN_METHOD_VI);                                   \
            if (mode == IT_SITES) {                                                                                       \
                if (kTab && use_table) emu_it_sweep<G, EpiSites, LK, M, kTab>(g, q.io, sa, cub, L, q.nc, q.spre.data(), q.sinit.data(), q.fs.data(), q.sw()); \
                else emu_it_sweep<G, EpiSites, LK, M, false>(g, q.io, sa, cub, L, q.nc, q.spre.data(), q.sinit.data(), q.fs.data(), q.sw()); \
            } else {                                                                                                      \
                if (kTab && use_table) emu_it_sweep<G, EpiEnergy, LK, M, kTab>(g, q.io, sa, cub, L, q.nc, q.spre.data(), q.sinit.data(), q.fs.data(), q.sw()); \
                else emu_it_sweep<G, EpiEnergy, LK, M, false>(g, q.io, sa, cub, L, q.nc, q.spre.data(), q.sinit.data(), q.fs.data(), q.sw()); \
            }                                                                                                             \
            done = true;                                                                                                  \
        }
        BN_FOR_EACH_ITER_SITE(X)
#undef X
        if (!done) return -3;
        for (double v : q.p1) s1 += v;
        for (double v : q.p2) s2 += v;
        const long long o = off[r];
        if (mode == IT_SITES) {
            host_from_tiled(q.n, L, q.nc, q.sy_t, sy + o);
            host_from_tiled(q.n, L, q.nc, q.sR_t, sR + o);
        } else {
            host_from_tiled(q.n, L, q.nc, q.pm_t, pm + o);
            host_from_tiled(q.n, L, q.nc, q.pc_t, pc + o);
        }
    }
    if (sums) { sums[0] = s1; sums[1] = s2; }
    return 0;
}

}  // namespace bn

using namespace bn;

#define EMU_MATERN(X) X(BN_MATERN12, 1) X(BN_MATERN32, 1) X(BN_MATERN32, 2) X(BN_MATERN52, 1) X(BN_MATERN52, 2) \
    X(BN_MATERN72, 1)
#define EMU_ARR(X) X(3, 1) X(4, 2) X(2, 2)

extern "C" int emu_kalman_filter(const bn_kernel_spec* k, int form, long long N, int L, int world, const double* dt,
                                 const double* y, const double* R, const unsigned char* mask, int return_predict,
                                 double* ell, double* means, double* covs) {
#define X(FAM, NC)                                                                                          \
    if (k->family == FAM && k->n_components == NC) {                                                        \
        auto make = [&](long long o) { MaternGen<FAM, NC> g; g.spec = *k; g.dt = dt + o; return g; };        \
        return emu_kf(make, 0, form, N, L, world, y, NC, R, mask, return_predict, ell, means, covs,         \
                      MaternGen<FAM, NC>::d);                                                               \
    }
    EMU_MATERN(X)
#undef X
    return -1;
}

extern "C" int emu_rts_smoother(const bn_kernel_spec* k, int form, long long N, int L, int world, const double* dt,
                                const double* fm, const double* fP, int return_full, double* means, double* covs,
                                double* gains) {
#define X(FAM, NC)                                                                                          \
    if (k->family == FAM && k->n_components == NC) {                                                        \
        auto make = [&](long long o) { MaternGen<FAM, NC> g; g.spec = *k; g.dt = dt + o; return g; };        \
        return emu_rts(make, form, N, L, world, fm, fP, return_full, means, covs, gains);                   \
    }
    EMU_MATERN(X)
#undef X
    return -1;
}

extern "C" int emu_kf_arrays(int form, long long N, int L, int d, int D, const double* As, const double* Qs,
                             const double* H, const double* ys, const double* Rs, const double* m0, const double* P0,
                             const unsigned char* masks, int return_predict, double* ell, double* fms, double* fPs) {
#define X(DD, OD)                                                                                           \
    if (d == DD && D == OD) {                                                                               \
        auto make = [&](long long o) { return ArrayGen<DD, OD>{As + o * DD * DD, Qs + o * DD * DD, H, m0, P0}; }; \
        return emu_kf(make, 0, form, N, L, 1, ys, OD, Rs, masks, return_predict, ell, fms, fPs, DD);        \
    }
    EMU_ARR(X)
#undef X
    return -1;
}

extern "C" int emu_rts_arrays(int form, long long N, int L, int d, int Df, const double* fms, const double* fPs,
                              const double* As, const double* Qs, const double* H, int return_full, double* sms,
                              double* sPs, double* gains) {
#define X(DD, OD)                                                                                           \
    if (d == DD && Df == OD) {                                                                              \
        auto make = [&](long long o) {                                                                      \
            return ArrayGen<DD, OD>{As + o * DD * DD, Qs + o * DD * DD, H, nullptr, nullptr};               \
        };                                                                                                  \
        return emu_rts(make, form, N, L, 1, fms, fPs, return_full, sms, sPs, gains);                        \
    }
    EMU_ARR(X)
#undef X
    return -1;
}

// ---------------------------------------------------------------------------- sites
// host SiteCtx: 1-D rule by value, probit table on the host (use_table), multi-latent rule = the host arrays
struct HostSite {
    Cub1 cub;
    SiteCtx sc;
    HostSite(const bn_site_args* a, int use_table) {
        const bool het = a->D == 2;
        make_cub1(het ? 0 : a->Q, het ? nullptr : a->cub_x, het ? nullptr : a->cub_w, cub);
        sc = SiteCtx{&cub, use_table ? probit_table_host().data() : nullptr, a->cub_x, a->cub_w};
    }
};

extern "C" double emu_probit_log_phi(double f) { return probit_log_phi(probit_table_host().data(), f); }

extern "C" int emu_site_update(const bn_site_args* a, int use_table) {
    HostSite hs(a, use_table);
    double s1 = 0.0, s2 = 0.0;
#define X(L, M)                                                   \
    if (a->likelihood == L && a->method == M) {                   \
        for (long long n = 0; n < a->N; ++n) {                    \
            double d1 = 0.0, d2 = 0.0;                            \
            if (use_table) site_update_step<L, M, true>(*a, hs.sc, n, d1, d2);     \
            else site_update_step<L, M, false>(*a, hs.sc, n, d1, d2);                \
            s1 += d1;                                             \
            s2 += d2;                                             \
        }                                                         \
        if (a->diffs) {                                           \
            a->diffs[0] = s1 / ((double)a->N * a->D);             \
            a->diffs[1] = s2 / ((double)a->N * a->D * a->D);      \
        }                                                         \
        return 0;                                                 \
    }
    BN_FOR_EACH_SITE(X)
#undef X
    return -1;
}

extern "C" int emu_expected_density(const bn_site_args* a, double* values, double* sum, int use_table) {
    HostSite hs(a, use_table);
    double s = 0.0;
#define X(L, M)                                                   \
    if (a->likelihood == L && a->method == M) {                   \
        for (long long n = 0; n < a->N; ++n) {                    \
            double v = use_table ? expected_density_step<L, M, true>(*a, hs.sc, n)    \
                                 : expected_density_step<L, M, false>(*a, hs.sc, n);        \
            if (values) values[n] = v;                            \
            if (!isnan(v)) s += v;                                \
        }                                                         \
        *sum = s;                                                 \
        return 0;                                                 \
    }
    BN_FOR_EACH_SITE(X)
#undef X
    return -1;
}

extern "C" int emu_gaussian_expected_log_lik(long long N, int D, const double* py, const double* pm, const double* pV,
                                             const double* pR, const unsigned char* mask, double* sum) {
    double s = 0.0;
    for (long long n = 0; n < N; ++n)
        s += (D == 1) ? gaussian_ell_step<1>(py, pm, pV, pR, mask, n) : gaussian_ell_step<2>(py, pm, pV, pR, mask, n);
    *sum = s;
    return 0;
}

extern "C" int emu_ep_pseudo_density(long long N, int D, double power, int with_const, const double* py,
                                     const double* pR, const double* pm, const double* pV, const double* n1,
                                     const double* n2, const unsigned char* mask, double* sum) {
    double s = 0.0;
    for (long long n = 0; n < N; ++n) {
        double v = (D == 1) ? ep_pseudo_step<1>(power, with_const, py, pR, pm, pV, n1, n2, mask, n)
                            : ep_pseudo_step<2>(power, with_const, py, pR, pm, pV, n1, n2, mask, n);
        if (!isnan(v)) s += v;
    }
    *sum = s;
    return 0;
}

extern "C" int emu_update_posterior(const bn_kernel_spec* k, long long N, int L, int world, const double* dt,
                                    const double* y, const double* R, const unsigned char* mask, double* ell,
                                    double* post_mean, double* post_cov, double* dvar, double* dlen) {
#define X(FAM, NC) \
    if (k->family == FAM && k->n_components == NC) \
        return emu_up<FastGen<FAM, NC>>(k, N, L, world, dt, y, R, mask, ell, post_mean, post_cov, dvar, dlen);
    EMU_MATERN(X)
#undef X
    return -1;
}

// ---------------------------------------------------------------------------- one emulated rank, phase by phase
// Host twin of bn_up_shard_{reduce,filter,smooth} for ONE rank (the carries are exchanged by the caller --
// tests/test_distributed_cpu.py does it with torch.distributed/gloo between real processes).
struct EmuRankBase {
    virtual ~EmuRankBase() {}
    virtual int kf_len() = 0;
    virtual int rts_len() = 0;
    virtual void reduce(const double* dt, const double* y, const double* R, double* carry) = 0;
    virtual void filter(const double* kf_carries, const unsigned char* mask, double* ell, double* rts_carry) = 0;
    virtual void smooth(const double* rts_carries, double* pm, double* pc, double* dvar, double* dlen) = 0;
};

template <class G>
struct EmuRank : EmuRankBase {
    static constexpr int d = G::d, D = G::D;
    using FA = FilterAlg<d>;
    using SA = SmootherAlg<d>;
    G g;
    bn_kernel_spec spec_;
    long long n, nc;
    int L, rank, world;
    UpIO io;
    std::vector<double> agg, fpre, sel, spre, fs, s0, sinit;
    EmuRank(const bn_kernel_spec* k, long long n_, int L_, int rank_, int world_) : n(n_), L(L_), rank(rank_), world(world_) {
        g.prepare(*k);
        spec_ = *k;
        nc = (n + L - 1) / L;
        agg.assign((size_t)nc * FA::kElem, 0.0);
        fpre = agg;
        sel.assign((size_t)nc * SA::kElem, 0.0);
        spre = sel;
        fs.assign((size_t)fs_doubles(nc, L, d + symn(d)), 0.0);
        s0.assign(64, 0.0);
        sinit.assign(64, 0.0);
    }
    int kf_len() override { return FA::kCarry; }
    int rts_len() override { return SA::kCarry; }
    void reduce(const double* dt, const double* y, const double* R, double* carry) override {
        io = UpIO{n, dt, y, R, nullptr, nullptr, nullptr};
        DirectCtx<D> cx{io};
        for (long long c = 0; c < nc; ++c) up_reduce_chunk(g, cx, n, L, nc, rank == 0, agg.data(), c, true);
        host_scan<FA>(agg.data(), nc, fpre.data());
        export_carry_body<FA>(fpre.data(), nc, carry);
    }
    void filter(const double* kf_carries, const unsigned char* mask, double* ell, double* rts_carry) override {
        io.mask = mask;
        DirectCtx<D> cx{io};
        fold_carries_body<FA>(kf_carries, 0, rank, 1, s0.data());
        std::vector<double> partials(nc, 0.0);
        for (long long c = 0; c < nc; ++c)
            up_filter_chunk<G, true>(g, cx, n, L, nc, rank == 0, fpre.data(), s0.data(), fs.data(), partials.data(), c, true);
        double tot = 0.0;
        for (double v : partials) tot += v;
        if (ell) *ell = tot;
        for (long long c = 0; c < nc; ++c) up_selem_chunk<G>(n, L, nc, rank != 0, agg.data(), s0.data(), fs.data(), sel.data(), c);
        host_scan<SA>(sel.data(), nc, spre.data());
        up_export_scarry<d>(spre.data(), nc, rank == world - 1, n, L, fs.data(), rts_carry);
    }
    void smooth(const double* rts_carries, double* pm, double* pc, double* dvar, double* dlen) override {
        io.post_mean = pm;
        io.post_cov = pc;
        DirectCtx<D> cx{io};
        if (rank != world - 1) fold_carries_body<SA>(rts_carries, world - 1, rank, -1, sinit.data());
        else up_last_state<d>(n, L, fs.data(), sinit.data());
        if (!dvar) {
            for (long long c = 0; c < nc; ++c)
                up_smooth_chunk<G, false>(g, cx, n, L, nc, spre.data(), sinit.data(), fs.data(), c, true);
            return;
        }
        constexpr int NF = GradAcc<G>::kFields;  // this rank's share of the hyper-gradient (summed over ranks by the caller)
        std::vector<double> gpart((size_t)nc * NF, 0.0), fields(NF, 0.0);
        for (long long c = 0; c < nc; ++c)
            up_smooth_chunk<G, true>(g, cx, n, L, nc, spre.data(), sinit.data(), fs.data(), c, true, rank == 0,
                                     s0.data(), gpart.data());
        for (int f = 0; f < NF; ++f)
            for (long long c = 0; c < nc; ++c) fields[f] += gpart[(size_t)f * nc + c];
        grad_finish<G>(spec_, fields.data(), dvar, dlen);
    }
};

extern "C" void* emu_rank_new(const bn_kernel_spec* k, long long n, int L, int rank, int world) {
    if (L % kUpTJ != 0) return nullptr;
#define X(FAM, NC) \
    if (k->family == FAM && k->n_components == NC) return new EmuRank<FastGen<FAM, NC>>(k, n, L, rank, world);
    EMU_MATERN(X)
#undef X
    return nullptr;
}
extern "C" void emu_rank_free(void* h) { delete (EmuRankBase*)h; }
extern "C" int emu_rank_kf_len(void* h) { return ((EmuRankBase*)h)->kf_len(); }
extern "C" int emu_rank_rts_len(void* h) { return ((EmuRankBase*)h)->rts_len(); }
extern "C" void emu_rank_reduce(void* h, const double* dt, const double* y, const double* R, double* carry) {
    ((EmuRankBase*)h)->reduce(dt, y, R, carry);
}
extern "C" void emu_rank_filter(void* h, const double* kf_carries, const unsigned char* mask, double* ell,
                                double* rts_carry) {
    ((EmuRankBase*)h)->filter(kf_carries, mask, ell, rts_carry);
}
extern "C" void emu_rank_smooth(void* h, const double* rts_carries, double* pm, double* pc, double* dvar, double* dlen) {
    ((EmuRankBase*)h)->smooth(rts_carries, pm, pc, dvar, dlen);
}

// fused iteration pass on tiled state (iter_impl.cuh); linear arrays in and out
extern "C" int emu_iter_pass(const bn_kernel_spec* k, long long N, int L, int world, const double* dt, const double* y,
                             double* sy, double* sR, const unsigned char* mask, int mode, int method, int lik,
                             double lik_param, int Q, const double* cx, const double* cw, double lr, double power,
                             int ensure_psd, int use_table, double* ell, double* sums, double* pm, double* pc, int spec,
                             double* jstar_mean) {
    if (jstar_mean) *jstar_mean = 0.0;
#define X(FAM)                                                                                                          \
    if (k->family == FAM && k->n_components == 1)                                                                       \
        return emu_it<FastGen<FAM, 1>>(k, N, L, world, dt, y, sy, sR, mask, mode, method, lik, lik_param, Q, cx, cw, lr, \
                                       power, ensure_psd, use_table, ell, sums, pm, pc, spec, jstar_mean);
    X(BN_MATERN12) X(BN_MATERN32) X(BN_MATERN52) X(BN_MATERN72)
#undef X
    return -1;
}

// ---------------------------------------------------------------------------- warp-cooperative small-d path (gd_impl.cuh)
// phase 2 as gd.cu drives it: levels of groups going up, one combine per element going down
template <bool FILTER>
static void emu_gd_scan(int d, long long n, const double* agg, double* prefix, std::vector<double>& smem) {
    const GdW w{0, 1};
    const int ne = gd_elem<FILTER>(d);
    std::vector<double> upper((size_t)gd_scan_upper_elems(n) * ne);
    std::vector<long long> cnt{n};
    std::vector<double*> arr{prefix};
    const double* in = agg;
    double* up = upper.data();
    while (true) {
        const long long m = cnt.back(), groups = (m + kGdScanGroup - 1) / kGdScanGroup;
        for (long long g = 0; g < groups; ++g)
            gd_scan_group<FILTER>(w, d, m, in, arr.back(), groups > 1 ? up : nullptr, g, GdPool{smem.data()});
        if (groups <= 1) break;
        cnt.push_back(groups); arr.push_back(up); in = up;
        up += groups * ne;
    }
    for (int l = (int)cnt.size() - 2; l >= 0; --l)
        for (long long i = kGdScanGroup; i < cnt[l]; ++i) gd_scan_down<FILTER>(w, d, arr[l], arr[l + 1], i, GdPool{smem.data()});
}

// the chunk bodies with (lane, lanes) = (0, 1): every lane-strided loop covers all entries, __syncwarp is a no-op
extern "C" int emu_gd_kf(int form, long long N, int L, int d, int D, const double* As, const double* Qs, const double* H,
                         const double* ys, const double* Rs, const double* m0, const double* P0, const unsigned char* masks,
                         int return_predict, double* ell, double* fms, double* fPs) {
    if (d < 1 || d > kGdMaxD || D < 1 || D > d) return -1;
    const GdW w{0, 1};
    std::vector<double> smem((size_t)gd_pool_doubles(d));
    GdKf a{N, d, D, As, Qs, H, ys, Rs, m0, P0, masks, return_predict, fms, fPs};
    if (form == BN_SEQUENTIAL) {
        GdPool pool{smem.data()};
        double* m = pool.take(d);
        double* P = pool.take(d * d);
        gd_copy(w, m, m0, d);
        gd_copy(w, P, P0, d * d);
        const double e = gd_kf_run(w, a, 0, N, m, P, ell != nullptr, pool);
        if (ell) *ell = e;
        return 0;
    }
    const long long nchunks = (N + L - 1) / L;
    std::vector<double> agg((size_t)nchunks * gd_felem(d)), prefix((size_t)nchunks * gd_felem(d));
    for (long long c = 0; c < nchunks; ++c) gd_kf_reduce_chunk(w, a, L, c, agg.data(), GdPool{smem.data()});
    emu_gd_scan<true>(d, nchunks, agg.data(), prefix.data(), smem);
    double tot = 0.0;
    for (long long c = 0; c < nchunks; ++c) tot += gd_kf_apply_chunk(w, a, L, c, prefix.data(), ell != nullptr, GdPool{smem.data()});
    if (ell) *ell = tot;
    return 0;
}

extern "C" int emu_gd_rts(int form, long long N, int L, int d, int Df, const double* fms, const double* fPs,
                          const double* As, const double* Qs, const double* H, int return_full, double* sms, double* sPs,
                          double* gains) {
    if (d < 1 || d > kGdMaxD || Df < 1 || Df > d) return -1;
    const GdW w{0, 1};
    std::vector<double> smem((size_t)gd_pool_doubles(d));
    GdRts a{N, d, Df, fms, fPs, As, Qs, H, return_full, sms, sPs, gains};
    if (form == BN_SEQUENTIAL) {
        GdPool pool{smem.data()};
        double* sm = pool.take(d);
        double* sP = pool.take(d * d);
        gd_fill(w, sm, 0.0, d);
        gd_fill(w, sP, 0.0, d * d);
        gd_rts_run(w, a, 0, N, sm, sP, true, pool);
        return 0;
    }
    const long long nchunks = (N + L - 1) / L;
    std::vector<double> agg((size_t)nchunks * gd_selem(d)), prefix((size_t)nchunks * gd_selem(d));
    for (long long c = 0; c < nchunks; ++c) gd_rts_reduce_chunk(w, a, L, nchunks, c, agg.data(), GdPool{smem.data()});
    emu_gd_scan<false>(d, nchunks, agg.data(), prefix.data(), smem);
    for (long long c = 0; c < nchunks; ++c) gd_rts_apply_chunk(w, a, L, nchunks, c, prefix.data(), GdPool{smem.data()});
    return 0;
}
