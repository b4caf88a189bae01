// C ABI of the fused inference iteration on chunk-tiled state (include/bn_b200.h: bn_iter_*).
// Kernels: iter_impl.cuh, instantiated per Matern family in iter_m12.cu .. iter_m72.cu.
// The same text builds the fp32 entry points (bn_iter_*_f32; iter32.cu defines BN_REAL32 and BN_NS = bn32): `real` is
// then a float, every array of the argument block is fp32, and ell / sums / carries are float.
#include <cstdlib>
#include <cmath>
#include <mutex>
#include "iter_impl.cuh"

#ifdef BN_REAL32
#define BN_ITER_FN(name) name##_f32
typedef float bn_abi_real;
#else
#define BN_ITER_FN(name) name
typedef double bn_abi_real;
#endif

namespace BN_NS {
// BN_B200_SPEC_FILTER=0 keeps phase 1 a pure reduction (A/B validation of the speculative filter pass)
static bool spec_filter_enabled() {
    static const bool on = [] {
        const char* e = getenv("BN_B200_SPEC_FILTER");
        return !(e && e[0] == '0');
    }();
    return on;
}

static int spec_env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

int it_group_m12(const ItCall&);
int it_group_m32(const ItCall&);
int it_group_m52(const ItCall&);
int it_group_m72(const ItCall&);
}  // namespace BN_NS
namespace bn { bool probit_table_enabled(); }
namespace BN_NS {
#ifdef BN_REAL32
// the fp32 table in device memory, filled once per device; handed to the kernels through the same (opaque) pointer
// the fp64 build uses for its packed table
__device__ __align__(16) float g_probit_tab32[kPt32N * 4];
static bool g_probit32_ready[64] = {false};
static std::mutex g_probit32_mutex;
int probit_table_device(cudaStream_t, const double** tab) {
    int dev = 0;
    BN_CUDA(cudaGetDevice(&dev));
    BN_REQUIRE(dev >= 0 && dev < 64, "device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_probit32_mutex);
    if (!g_probit32_ready[dev]) {  // synchronous: visible to every stream when this returns
        BN_CUDA(cudaMemcpyToSymbol(g_probit_tab32, probit_table32_host().data(), kPt32Bytes, 0, cudaMemcpyHostToDevice));
        g_probit32_ready[dev] = true;
    }
    void* p = nullptr;
    BN_CUDA(cudaGetSymbolAddress(&p, g_probit_tab32));
    *tab = (const double*)p;
    return 0;
}
#endif

static int it_dispatch(const ItCall& c) {
    int r;
    if ((r = it_group_m12(c)) != kNotHandled) return r;
    if ((r = it_group_m32(c)) != kNotHandled) return r;
    if ((r = it_group_m52(c)) != kNotHandled) return r;
    if ((r = it_group_m72(c)) != kNotHandled) return r;
    set_error("the fused iteration supports one Matern component, got family %d with %d components", c.spec->family,
              c.spec->n_components);
    return -1;
}

static int it_check_spec(const bn_kernel_spec* k, int64_t N) {
    BN_REQUIRE(k != nullptr, "kernel spec is null");
    BN_REQUIRE(k->n_components == 1 && family_dim(k->family) > 0,
               "the fused iteration supports one Matern component, got family %d with %d components", k->family, k->n_components);
    BN_REQUIRE(N > 0, "N must be positive");
    return 0;
}

static size_t it_ws_bytes(const bn_kernel_spec* k, int64_t N) {
    switch (family_dim(k->family)) {
        case 1: return (it_ws_doubles<1>(N) + 64) * sizeof(real);
        case 2: return (it_ws_doubles<2>(N) + 64) * sizeof(real);
        case 3: return (it_ws_doubles<3>(N) + 64) * sizeof(real);
        case 4: return (it_ws_doubles<4>(N) + 64) * sizeof(real);
    }
    return 0;
}

// fills an ItCall from the public argument block; the 1-D rule goes by value into `cub`
static int it_make_call(const bn_kernel_spec* k, const bn_iter_args* a, int mode, int phase, Cub1& cub, ItCall& c) {
    BN_REQUIRE(a != nullptr, "iteration args are null");
    if (int rc = it_check_spec(k, a->N)) return rc;
    BN_REQUIRE(a->rank >= 0 && a->rank < a->world, "rank %d outside world %d", a->rank, a->world);
    BN_REQUIRE(mode == BN_ITER_PLAIN || mode == BN_ITER_SITES || mode == BN_ITER_ENERGY, "unknown mode %d", mode);
    BN_REQUIRE(a->dt_t != nullptr, "dt_t is null");
    const bool front = (phase == UP_ALL || phase == UP_REDUCE || phase == UP_FILTER);
    const bool back = (phase == UP_ALL || phase == UP_SMOOTH);
    if (front || (back && mode != BN_ITER_PLAIN)) BN_REQUIRE(a->site_mean_t && a->site_cov_t, "tiled site arrays are null");
    BN_REQUIRE((a->post_mean == nullptr) == (a->post_cov == nullptr), "give both linear posterior arrays or neither");
    if (back && mode != BN_ITER_SITES)
        BN_REQUIRE((a->post_mean_t && a->post_cov_t) || a->post_mean, "posterior output arrays are null");
    c = ItCall{};
    c.spec = k;
    c.io = ItIO{a->N, (const real*)a->dt_t, (const real*)a->y_t, (real*)a->site_mean_t, (real*)a->site_cov_t, a->mask_t,
                (real*)a->post_mean_t, (real*)a->post_cov_t, (real*)a->post_mean, (real*)a->post_cov};
    c.mode = mode;
    c.method = a->method;
    c.likelihood = a->likelihood;
#ifdef BN_REAL32
    c.use_table = ::bn::probit_table_enabled() ? 1 : 0;  // the fp32 build has its own 8 KB table (probit_table32.cuh)
#else
    c.use_table = ::bn::probit_table_enabled() ? 1 : 0;
#endif
    c.sa = ItSiteArgs{a->lik_param, a->lr, a->power, a->ensure_psd, 0, nullptr, nullptr};
    c.cub = &cub;
    c.phase = phase;
    c.rank = a->rank;
    c.world = a->world;
    c.spec_filter = spec_filter_enabled() ? 1 : 0;
    c.spec_min_chunk = spec_env_int("BN_B200_SPEC_MIN_CHUNK", 0);
    {
        const int lg = spec_env_int("BN_B200_SPEC_LOG2", 0);  // threshold 2^-lg (tuning aid; 0 = the built-in one)
        c.spec_thr = lg > 0 ? ldexp(1.0, -lg) : 0.0;
    }
    c.want_ell = a->want_ell;
    if (back && mode != BN_ITER_PLAIN) {
        BN_REQUIRE(a->y_t != nullptr, "y_t is null");
        BN_REQUIRE(a->method == BN_METHOD_VI || a->method == BN_METHOD_NEWTON || a->method == BN_METHOD_EP,
                   "the fused epilogues cover VI, Newton and EP, got method %d", a->method);
        if (a->method == BN_METHOD_EP) BN_REQUIRE(a->power > 0.0, "EP power must be positive");
        const bool closed = a->method == BN_METHOD_NEWTON ||
                            (a->method == BN_METHOD_VI && (a->likelihood == BN_LIK_GAUSSIAN || a->likelihood == BN_LIK_POISSON_EXP)) ||
                            (a->method == BN_METHOD_EP && a->likelihood == BN_LIK_GAUSSIAN);
        if (!closed) BN_REQUIRE(a->Q > 0 && a->Q <= kMaxQ1 && a->cub_x_host && a->cub_w_host, "cubature rule missing or larger than %d points", kMaxQ1);
        if (a->likelihood == BN_LIK_GAUSSIAN || a->likelihood == BN_LIK_POISSON_EXP) BN_REQUIRE(a->lik_param > 0.0, "likelihood parameter must be positive");
        const bool has = a->Q > 0 && a->Q <= kMaxQ1 && a->cub_x_host && a->cub_w_host;
        make_cub1(has ? a->Q : 0, has ? a->cub_x_host : nullptr, has ? a->cub_w_host : nullptr, cub);
    } else {
        make_cub1(0, nullptr, nullptr, cub);
    }
    return 0;
}

template <typename T>
static int it_transpose(const bn_kernel_spec* k, int64_t N, const T* in, T* out, T fill, bool to_tiled, cudaStream_t st) {
    if (int rc = it_check_spec(k, N)) return rc;
    BN_REQUIRE(in && out, "null array");
    const ChunkPlan cp = up_plan_chunks(N, false, family_dim(k->family));
    const long long tiles = (cp.nchunks + 31) / 32;
    dim3 grid((unsigned)tiles, (unsigned)((cp.L + 31) / 32));
    if (to_tiled) {
        BN_LAUNCH("it_to_tiled", st, (it_transpose_kernel<T, true><<<grid, 256, 0, st>>>(N, cp.L, cp.nchunks, in, out, fill)));
        const long long body = tiles * 32LL * cp.L, total = tl_len(cp.nchunks, cp.L);
        it_fill_pad_kernel<T><<<(unsigned)((total - body + 255) / 256), 256, 0, st>>>(out, body, total, fill);
    } else {
        BN_LAUNCH("it_from_tiled", st, (it_transpose_kernel<T, false><<<grid, 256, 0, st>>>(N, cp.L, cp.nchunks, in, out, fill)));
    }
    BN_CUDA(cudaGetLastError());
    return 0;
}
}  // namespace BN_NS

using namespace BN_NS;

#ifndef BN_REAL32  // the chunk plan does not depend on the scalar type
extern "C" int bn_iter_chunk_len(const bn_kernel_spec* k, int64_t N) {
    if (it_check_spec(k, N)) return -1;
    return up_plan_chunks(N, false, family_dim(k->family)).L;
}

extern "C" int64_t bn_iter_tiled_len(const bn_kernel_spec* k, int64_t N) {
    if (it_check_spec(k, N)) return -1;
    const ChunkPlan cp = up_plan_chunks(N, false, family_dim(k->family));
    return tl_len(cp.nchunks, cp.L);
}

extern "C" int bn_iter_to_tiled_u8(const bn_kernel_spec* k, int64_t N, const uint8_t* x, uint8_t* x_t, void* stream) {
    return it_transpose<unsigned char>(k, N, x, x_t, (unsigned char)0, true, (cudaStream_t)stream);
}
#endif

extern "C" size_t BN_ITER_FN(bn_iter_workspace_bytes)(const bn_kernel_spec* k, int64_t N) {
    if (it_check_spec(k, N)) return 0;
    return it_ws_bytes(k, N);
}

extern "C" int BN_ITER_FN(bn_iter_to_tiled)(const bn_kernel_spec* k, int64_t N, const bn_abi_real* x, bn_abi_real* x_t,
                                            bn_abi_real fill, void* stream) {
    return it_transpose<bn_abi_real>(k, N, x, x_t, fill, true, (cudaStream_t)stream);
}

extern "C" int BN_ITER_FN(bn_iter_from_tiled)(const bn_kernel_spec* k, int64_t N, const bn_abi_real* x_t, bn_abi_real* x,
                                              void* stream) {
    return it_transpose<bn_abi_real>(k, N, x_t, x, (bn_abi_real)0, false, (cudaStream_t)stream);
}

extern "C" int BN_ITER_FN(bn_iter_pass)(const bn_kernel_spec* k, const bn_iter_args* a, int mode, bn_abi_real* ell,
                                        bn_abi_real* sums, void* workspace, size_t workspace_bytes, void* stream) {
    Cub1 cub;
    ItCall c;
    if (int rc = it_make_call(k, a, mode, UP_ALL, cub, c)) return rc;
    BN_REQUIRE(a->world == 1 && a->rank == 0, "bn_iter_pass runs one shard; use the bn_iter_shard_* phases for world %d", a->world);
    c.ell = (real*)ell;
    c.want_ell = (ell != nullptr);
    c.sums = (real*)sums;
    c.ws = workspace;
    c.ws_bytes = workspace_bytes;
    c.st = (cudaStream_t)stream;
    return it_dispatch(c);
}

extern "C" int BN_ITER_FN(bn_iter_shard_reduce)(const bn_kernel_spec* k, const bn_iter_args* a, bn_abi_real* kf_carry,
                                                void* workspace, size_t workspace_bytes, void* stream) {
    Cub1 cub;
    ItCall c;
    if (int rc = it_make_call(k, a, BN_ITER_PLAIN, UP_REDUCE, cub, c)) return rc;
    BN_REQUIRE(kf_carry != nullptr, "carry output is null");
    c.carry_out = (real*)kf_carry;
    c.ws = workspace;
    c.ws_bytes = workspace_bytes;
    c.st = (cudaStream_t)stream;
    return it_dispatch(c);
}

extern "C" int BN_ITER_FN(bn_iter_shard_filter)(const bn_kernel_spec* k, const bn_iter_args* a, const bn_abi_real* kf_carries,
                                                bn_abi_real* ell, bn_abi_real* rts_carry, void* workspace,
                                                size_t workspace_bytes, void* stream) {
    Cub1 cub;
    ItCall c;
    if (int rc = it_make_call(k, a, BN_ITER_PLAIN, UP_FILTER, cub, c)) return rc;
    BN_REQUIRE(kf_carries && rts_carry, "null carry array");
    BN_REQUIRE((ell != nullptr) == (a->want_ell != 0), "want_ell of the argument block must say whether ell is requested");
    c.carries = (const real*)kf_carries;
    c.carry_out = (real*)rts_carry;
    c.ell = (real*)ell;
    c.ws = workspace;
    c.ws_bytes = workspace_bytes;
    c.st = (cudaStream_t)stream;
    return it_dispatch(c);
}

extern "C" int BN_ITER_FN(bn_iter_shard_smooth)(const bn_kernel_spec* k, const bn_iter_args* a, int mode,
                                                const bn_abi_real* rts_carries, bn_abi_real* sums, void* workspace,
                                                size_t workspace_bytes, void* stream) {
    Cub1 cub;
    ItCall c;
    if (int rc = it_make_call(k, a, mode, UP_SMOOTH, cub, c)) return rc;
    BN_REQUIRE(rts_carries != nullptr, "null carry array");
    c.carries = (const real*)rts_carries;
    c.sums = (real*)sums;
    c.ws = workspace;
    c.ws_bytes = workspace_bytes;
    c.st = (cudaStream_t)stream;
    return it_dispatch(c);
}
