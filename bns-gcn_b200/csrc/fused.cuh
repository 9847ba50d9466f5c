// fused.cuh -- the "tail" of the epoch as a handful of kernels (included by bnsgcn.cu, same translation unit).
//
// Round 1 left ~59 % of the 8-GPU epoch in small launches that are neither SpMM nor GEMM: the
// loss (index + log_softmax + nll + their backward), Adam (15 multi-tensor launches), per-peer put / wait / scatter-add
// launches of the boundary exchange (7 + 7 + 7 per layer per direction at 8 partitions), per-peer slot-map updates,
// per-parameter gradient scaling, weight transposes and pads.  Each group becomes ONE launch here:
//
//   xent_kernel            loss + d(logits) of train.py:406-408 (CrossEntropyLoss / BCEWithLogitsLoss, reduction='sum',
//                          over the train rows), deterministic block-ordered reduction
//   adam_kernel            torch.optim.Adam over ONE flat parameter arena (train.py:362, :413)
//   derive_kernel          refreshes the cached W^T / bias sums of the arena after a step (table driven)
//   p2p_put_all_kernel     pack + store the rows of ALL peers into their slabs, then publish all flags
//   p2p_wait_all_kernel    one warp waits for all peers' flags
//   scatter_rows_all_kernel  G[selected_j] += R_j / ratio_j for all peers j in the reference's order, race-free by
//                          walking destination rows through per-peer inverse maps
//   epoch_maps_kernel      slot map (construct_graph, train.py:256-281) and inverse maps of all peers
//   p2p_put_ids_kernel     sampled id lists of all peers into their slabs (data_transfer NODE, train.py:389)
//   compact_cols_kernel    per-epoch compaction of the halo matrix to the sampled columns (chunk-local, no scan)
//   dropout / scale_rows   small element-wise helpers with the Philox mask of ln_relu_dropout_kernel

namespace {

constexpr int kMaxPeers = BNS_MAX_PEERS;

// =====================================================================================================================
// loss
// =====================================================================================================================
struct XentArgs {
    const float *logits; int64_t ld; int64_t n; int32_t C;
    const int64_t *labels;               // CE: class index per row
    const float *labels_f; int64_t ldl;  // BCE: target per (row, class)
    const uint8_t *mask;                 // train mask (bool bytes) or NULL = every row
    float grad_scale;
    float *dlogits; int64_t ldd; int32_t C_out;
    float *partial; unsigned int *ticket; float *loss_out;
};

template <bool BCE>
__global__ void __launch_bounds__(kThreads) xent_kernel(XentArgs a) {
    __shared__ float s_w[kWarps];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * kWarps;
    float wloss = 0.f;      // meaningful on lane 0
    for (int64_t row = (int64_t)blockIdx.x * kWarps + w; row < a.n; row += warps_total) {
        const float *x = a.logits + row * a.ld;
        float *d = a.dlogits + row * a.ldd;
        const bool on = a.mask == nullptr || a.mask[row] != 0;
        if (!on) {
            for (int c = lane; c < a.C_out; c += 32) d[c] = 0.f;
            continue;
        }
        if (!BCE) {
            float m = -INFINITY;
            for (int c = lane; c < a.C; c += 32) m = fmaxf(m, x[c]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float s = 0.f;
            for (int c = lane; c < a.C; c += 32) s += expf(x[c] - m);
            s = warp_sum(s);
            const int64_t lab = a.labels[row];
            const float inv = 1.f / s;
            for (int c = lane; c < a.C_out; c += 32) {
                float g = 0.f;
                if (c < a.C) g = (expf(x[c] - m) * inv - (c == lab ? 1.f : 0.f)) * a.grad_scale;
                d[c] = g;
            }
            if (lane == 0) wloss += (m + logf(s)) - x[lab];
        } else {
            const float *y = a.labels_f + row * a.ldl;
            float l = 0.f;
            for (int c = lane; c < a.C_out; c += 32) {
                float g = 0.f;
                if (c < a.C) {
                    const float xv = x[c], yv = y[c];
                    l += fmaxf(xv, 0.f) - xv * yv + log1pf(expf(-fabsf(xv)));
                    g = (1.f / (1.f + expf(-xv)) - yv) * a.grad_scale;
                }
                d[c] = g;
            }
            l = warp_sum(l);
            if (lane == 0) wloss += l;
        }
    }
    if (lane == 0) s_w[w] = wloss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < kWarps; ++i) t += s_w[i];          // fixed order
        a.partial[blockIdx.x] = t;
        __threadfence();
        const unsigned int done = atomicAdd(a.ticket, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && w == 0) {                                   // last block: sum the per-block partials in a fixed tree
        __threadfence();
        float t = 0.f;
        for (unsigned i = lane; i < gridDim.x; i += 32) t += ((volatile float *)a.partial)[i];
        t = warp_sum(t);
        if (lane == 0) {
            *a.loss_out = t;
            *a.ticket = 0;                                    // re-armed for the next (stream-ordered) launch
        }
    }
}

inline unsigned xent_grid(int64_t n) {
    int64_t want = (n + kWarps - 1) / kWarps, cap = (int64_t)sm_count() * 4;
    return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

namespace { constexpr unsigned kXentMaxGrid = 2048; }

// per-block partial losses + the ticket counter; the caller zeroes it ONCE (the kernel re-arms the ticket itself)
extern "C" size_t bns_xent_workspace_bytes(void) { return (size_t)kXentMaxGrid * sizeof(float) + 256; }

extern "C" int bns_xent_f32(const float *logits, int64_t ld, int64_t n_rows, int32_t n_class, const int64_t *labels,
                            const float *labels_f, int64_t ldl, const uint8_t *mask, float grad_scale, float *loss_out,
                            float *dlogits, int64_t ldd, int32_t n_cols_out, void *ws, size_t ws_bytes, void *stream) {
    BNS_REQUIRE(n_rows >= 0 && n_class > 0 && n_cols_out >= n_class, "bns_xent_f32: bad shape");
    BNS_REQUIRE(logits && loss_out && dlogits && (labels || labels_f), "bns_xent_f32: NULL pointer");
    BNS_REQUIRE(ld >= n_class && ldd >= n_cols_out && (!labels_f || ldl >= n_class), "bns_xent_f32: bad leading dimension");
    BNS_REQUIRE(!(labels && labels_f), "bns_xent_f32: give class indices OR per-class targets");
    unsigned grid = xent_grid(n_rows);
    if (grid > kXentMaxGrid) grid = kXentMaxGrid;
    if (!ws || ws_bytes < bns_xent_workspace_bytes())
        return fail(BNS_E_WORKSPACE, "bns_xent_f32: workspace %zu bytes < %zu needed", ws_bytes, bns_xent_workspace_bytes());
    XentArgs a;
    a.logits = logits; a.ld = ld; a.n = n_rows; a.C = n_class; a.labels = labels; a.labels_f = labels_f; a.ldl = ldl;
    a.mask = mask; a.grad_scale = grad_scale; a.dlogits = dlogits; a.ldd = ldd; a.C_out = n_cols_out;
    a.partial = reinterpret_cast<float *>(ws);
    a.ticket = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(ws) + kXentMaxGrid * sizeof(float));   // zeroed by the caller once
    a.loss_out = loss_out;
    if (labels) xent_kernel<false><<<grid, kThreads, 0, as_stream(stream)>>>(a);
    else xent_kernel<true><<<grid, kThreads, 0, as_stream(stream)>>>(a);
    ++g_launches;
    BNS_CUDA(cudaGetLastError());
    return BNS_OK;
}

// =====================================================================================================================
// Adam over a flat arena + derived parameters
// =====================================================================================================================
namespace {

__global__ void __launch_bounds__(256) adam_kernel(float4 *__restrict__ p, const float4 *__restrict__ g, float4 *__restrict__ m,
                                                   float4 *__restrict__ v, int64_t n4, float lr, float beta1, float beta2,
                                                   float eps, float wd, const int64_t *__restrict__ step_dev) {
    const double step = (double)(*step_dev + 1);
    const float bc1 = (float)(1.0 - pow((double)beta1, step));
    const float sqrt_bc2 = (float)sqrt(1.0 - pow((double)beta2, step));
    const float step_size = lr / bc1;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
        float *pf = &pp.x, *gf = &gg.x, *mf = &mm.x, *vf = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float gr = gf[k];
            if (wd != 0.f) gr = fmaf(wd, pf[k], gr);
            mf[k] = mf[k] + (1.f - beta1) * (gr - mf[k]);                 // exp_avg.lerp_(grad, 1 - beta1)
            vf[k] = vf[k] * beta2 + (1.f - beta2) * gr * gr;              // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
            const float denom = sqrtf(vf[k]) / sqrt_bc2 + eps;
            pf[k] = pf[k] - step_size * (mf[k] / denom);
        }
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

}  // namespace

// One derived-parameter refresh (bns_derive_entry, bnsgcn.h): op 0 transpose, op 1 sum of two vectors

namespace {

__global__ void __launch_bounds__(256) derive_kernel(const bns_derive_entry *__restrict__ table, int64_t *step_dev) {
    __shared__ float tile[32][33];
    const bns_derive_entry e = table[blockIdx.y];
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && step_dev) *step_dev += 1;
    if (e.op == 1) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < e.rows; i += gridDim.x * blockDim.x) e.dst[i] = e.a[i] + e.b[i];
        return;
    }
    const int tiles_c = (e.cols + 31) / 32, tiles_r = (e.rows + 31) / 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 32 x 8
    for (int t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
        const int r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
        for (int j = ty; j < 32; j += 8) {
            const int r = r0 + j, c = c0 + tx;
            tile[j][tx] = (r < e.rows && c < e.cols) ? e.a[(int64_t)r * e.ld_a + c] : 0.f;
        }
        __syncthreads();
        for (int j = ty; j < 32; j += 8) {
            const int c = c0 + j, r = r0 + tx;
            if (c < e.cols && r < e.rows) e.dst[(int64_t)c * e.ld_dst + r] = tile[tx][j];
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" size_t bns_derive_entry_bytes(void) { return sizeof(bns_derive_entry); }

// step_dev: int64 on the device, the number of steps taken so far (the kernel uses step_dev + 1 and does NOT change it;
// bns_derive_refresh advances it -- call it after every step, with n_entries == 0 if nothing is derived)
extern "C" int bns_adam_step_f32(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, const int64_t *step_dev,
                                 void *stream) {
    BNS_REQUIRE(n >= 0 && n % 4 == 0, "bns_adam_step_f32: the arena length must be a multiple of 4");
    if (n == 0) return BNS_OK;
    BNS_REQUIRE(param && grad && exp_avg && exp_avg_sq && step_dev, "bns_adam_step_f32: NULL pointer");
    BNS_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                  reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15u) == 0, "bns_adam_step_f32: 16-byte alignment required");
    const int64_t n4 = n / 4;
    int64_t want = (n4 + 255) / 256, cap = (int64_t)sm_count() * 8;
    adam_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<float4 *>(param), reinterpret_cast<const float4 *>(grad), reinterpret_cast<float4 *>(exp_avg),
        reinterpret_cast<float4 *>(exp_avg_sq), n4, lr, beta1, beta2, eps, weight_decay, step_dev);
    ++g_launches;
    BNS_CUDA(cudaGetLastError());
    return BNS_OK;
}

namespace {
__global__ void step_inc_kernel(int64_t *step_dev) { *step_dev += 1; }
}  // namespace

// Refresh the derived parameters (table of bns_derive_entry on the device) and advance the step counter by one; with
// n_entries == 0 only the counter moves.  Enqueue it right after bns_adam_step_f32 on the same stream.
extern "C" int bns_derive_refresh(const void *table_dev, int32_t n_entries, int64_t *step_dev, void *stream) {
    BNS_REQUIRE(n_entries >= 0 && (n_entries == 0 || table_dev), "bns_derive_refresh: bad table");
    if (n_entries == 0) {
        if (!step_dev) return BNS_OK;
        step_inc_kernel<<<1, 1, 0, as_stream(stream)>>>(step_dev);
    } else {
        derive_kernel<<<dim3(32, (unsigned)n_entries), 256, 0, as_stream(stream)>>>(
            reinterpret_cast<const bns_derive_entry *>(table_dev), step_dev);
    }
    ++g_launches;
    BNS_CUDA(cudaGetLastError());
    return BNS_OK;
}

// =====================================================================================================================
// element-wise helpers
// =====================================================================================================================
namespace {

__global__ void __launch_bounds__(kThreads) dropout_kernel(const float *__restrict__ x, int64_t ldx, int64_t n, int32_t F,
                                                           float p, float keep_scale, uint64_t seed, uint64_t offset,
                                                           const uint64_t *__restrict__ offset_dev, float *__restrict__ y,
                                                           int64_t ldy) {
    if (offset_dev) offset += *offset_dev;
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * kWarps;
    for (int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); row < n; row += warps_total) {
        for (int vec = lane; vec * 4 < F; vec += 32) {
            const float4 v = *reinterpret_cast<const float4 *>(x + row * ldx + vec * 4);
            bool keep[4];
            drop_mask4(seed, offset, row, vec, p, keep);
            float4 o;
            o.x = keep[0] ? v.x * keep_scale : 0.f; o.y = keep[1] ? v.y * keep_scale : 0.f;
            o.z = keep[2] ? v.z * keep_scale : 0.f; o.w = keep[3] ? v.w * keep_scale : 0.f;
            *reinterpret_cast<float4 *>(y + row * ldy + vec * 4) = o;
        }
    }
}

__global__ void __launch_bounds__(kThreads) scale_rows_kernel(const float *__restrict__ x, int64_t ldx, int64_t n, int32_t F,
                                                              const float *__restrict__ rs, const float *__restrict__ bias,
                                                              float *__restrict__ y, int64_t ldy) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * kWarps;
    for (int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); row < n; row += warps_total) {
        const float s = rs ? rs[row] : 1.f;
        for (int f = lane * 4; f < F; f += 128) {
            float4 v = *reinterpret_cast<const float4 *>(x + row * ldx + f);
            v.x *= s; v.y *= s; v.z *= s; v.w *= s;
            if (bias) {
                const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + f));
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            *reinterpret_cast<float4 *>(y + row * ldy + f) = v;
        }
    }
}

}  // namespace

extern "C" int bns_dropout_f32(const float *x, int64_t ldx, int64_t n, int64_t F, float p, uint64_t seed, uint64_t offset,
                               const uint64_t *offset_dev, float *y, int64_t ldy, void *stream) {
    BNS_REQUIRE(n >= 0 && F > 0 && F % 4 == 0, "bns_dropout_f32: need F %% 4 == 0");
    if (n == 0) return BNS_OK;
    BNS_REQUIRE(x && y && ldx >= F && ldy >= F && ldx % 4 == 0 && ldy % 4 == 0, "bns_dropout_f32: bad pointer / leading dimension");
    BNS_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0, "bns_dropout_f32: unaligned");
    BNS_REQUIRE(p >= 0.f && p < 1.f, "bns_dropout_f32: p must be in [0, 1)");
    dropout_kernel<<<ln_grid(n), kThreads, 0, as_stream(stream)>>>(x, ldx, n, (int32_t)F, p, 1.f / (1.f - p), seed, offset,
                                                                   offset_dev, y, ldy);
    ++g_launches;
    BNS_CUDA(cudaGetLastError());
    return BNS_OK;
}

extern "C" int bns_scale_rows_f32(const float *x, int64_t ldx, int64_t n, int64_t F, const float *row_scale, const float *bias,
                                  float *y, int64_t ldy, void *stream) {
    BNS_REQUIRE(n >= 0 && F > 0 && F % 4 == 0, "bns_scale_rows_f32: need F %% 4 == 0");
    if (n == 0) return BNS_OK;
    BNS_REQUIRE(x && y && ldx >= F && ldy >= F && ldx % 4 == 0 && ldy % 4 == 0, "bns_scale_rows_f32: bad argument");
    BNS_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias)) & 15u) == 0,
                "bns_scale_rows_f32: unaligned");
    scale_rows_kernel<<<ln_grid(n), kThreads, 0, as_stream(stream)>>>(x, ldx, n, (int32_t)F, row_scale, bias, y, ldy);
    ++g_launches;
    BNS_CUDA(cudaGetLastError());
    return BNS_OK;
}

// =====================================================================================================================
// per-epoch maps: slot map (halo node -> slab row) and per-peer inverse maps (inner node -> position in selected_j)
// =====================================================================================================================
namespace {

__global__ void epoch_maps_kernel(bns_epoch_maps a) {
    const int64_t sel_total = a.sel_begin[a.n_seg], hop_total = a.hop_begin[a.n_seg];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < sel_total + hop_total;
         i += (int64_t)gridDim.x * blockDim.x) {
        if (i < sel_total) {
            int s = 0;
            while (s + 1 < a.n_seg && a.sel_begin[s + 1] <= i) ++s;
            if (a.inv[s]) a.inv[s][a.selected_cat[i]] = (int32_t)(i - a.sel_begin[s]);
        } else {
            const int64_t k = i - sel_total;
            int s = 0;
            while (s + 1 < a.n_seg && a.hop_begin[s + 1] <= k) ++s;
            const int64_t local = a.pos[s][a.one_hops_cat[k]];
            if (local >= a.n_in) a.slot[local - a.n_in] = (int32_t)k;           // U-numbering minus n_in == position in the cat
        }
    }
}

}  // namespace

// `fill_base` .. + fill_bytes: the slot map and the inverse maps live in ONE allocation; it is set to -1 (0xFF bytes)
// first, then the sampled entries are written.  Replaces the per-peer loop of train.py:256-281 and builds the inverse
// maps bns_scatter_rows_all_f32 walks.
extern "C" int bns_epoch_maps_update(const bns_epoch_maps *maps, void *fill_base, size_t fill_bytes, void *stream) {
    BNS_REQUIRE(maps && maps->n_seg >= 0 && maps->n_seg <= kMaxPeers, "bns_epoch_maps_update: bad segment table");
    cudaStream_t st = as_stream(stream);
    if (fill_base && fill_bytes) BNS_CUDA(cudaMemsetAsync(fill_base, 0xFF, fill_bytes, st));
    const int64_t total = maps->sel_begin[maps->n_seg] + maps->hop_begin[maps->n_seg];
    if (total == 0 || maps->n_seg == 0) return BNS_OK;
    epoch_maps_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(*maps);
    ++g_launches;
    BNS_CUDA(cudaGetLastError());
    return BNS_OK;
}

// =====================================================================================================================
// consolidated peer-mapped exchange
// =====================================================================================================================
namespace {

struct PutAllDev {
    int32_t n_seg;
    int64_t row_begin[kMaxPeers + 1];
    float *remote[kMaxPeers];
    unsigned long long *flag[kMaxPeers];
    int64_t src_begin[kMaxPeers];
    float div[kMaxPeers];
    const float *H; int64_t ldh; int32_t F;
    const int64_t *idx;                         // concatenated in segment order, or NULL
    int64_t ld_remote;
    unsigned long long flag_value; const unsigned long long *flag_value_dev; unsigned int *ticket;
};

template <bool VEC>
__global__ void __launch_bounds__(kThreads) p2p_put_all_kernel(PutAllDev a) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * kWarps, total = a.row_begin[a.n_seg];
    for (int64_t i = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); i < total; i += warps_total) {
        int s = 0;
        while (s + 1 < a.n_seg && a.row_begin[s + 1] <= i) ++s;
        const int64_t local = i - a.row_begin[s];
        const int64_t r = a.idx ? a.idx[i] : a.src_begin[s] + local;
        const float *src = a.H + r * a.ldh;
        float *d = a.remote[s] + local * a.ld_remote;
        const float div = a.div[s];
        if (VEC) {
            for (int f = lane * 4; f < a.F; f += 128) {
                float4 v = *reinterpret_cast<const float4 *>(src + f);
                v.x = __fdiv_rn(v.x, div); v.y = __fdiv_rn(v.y, div); v.z = __fdiv_rn(v.z, div); v.w = __fdiv_rn(v.w, div);
                *reinterpret_cast<float4 *>(d + f) = v;
            }
        } else {
            for (int f = lane; f < a.F; f += 32) d[f] = __fdiv_rn(src[f], div);
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(a.ticket, 1u);
        s_last = (done == gridDim.x - 1);
        if (s_last) atomicExch(a.ticket, 0u);          // re-armed: launches that share a ticket are stream-ordered
    }
    __syncthreads();
    if (s_last && (int)threadIdx.x < a.n_seg) {
        __threadfence_system();
        st_release_sys(a.flag[threadIdx.x], a.flag_value + (a.flag_value_dev ? *a.flag_value_dev : 0ull));
    }
}

struct PutIdsDev {
    int32_t n_seg;
    int64_t begin[kMaxPeers + 1];
    int64_t *remote[kMaxPeers];
    unsigned long long *flag[kMaxPeers];
    const int64_t *src;
    unsigned long long flag_value; const unsigned long long *flag_value_dev; unsigned int *ticket;
};

__global__ void __launch_bounds__(256) p2p_put_ids_kernel(PutIdsDev a) {
    const int64_t total = a.begin[a.n_seg];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int s = 0;
        while (s + 1 < a.n_seg && a.begin[s + 1] <= i) ++s;
        a.remote[s][i - a.begin[s]] = a.src[i];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(a.ticket, 1u);
        s_last = (done == gridDim.x - 1);
        if (s_last) atomicExch(a.ticket, 0u);
    }
    __syncthreads();
    if (s_last && (int)threadIdx.x < a.n_seg) {
        __threadfence_system();
        st_release_sys(a.flag[threadIdx.x], a.flag_value + (a.flag_value_dev ? *a.flag_value_dev : 0ull));
    }
}

struct WaitAllDev {
    int32_t n;
    const unsigned long long *flag[kMaxPeers];
};

__global__ void p2p_wait_all_kernel(WaitAllDev a, unsigned long long value, const unsigned long long *value_dev,
                                    unsigned long long timeout_ns) {
    if ((int)threadIdx.x >= a.n) return;
    if (value_dev) value += *value_dev;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(a.flag[threadIdx.x]) < value) {
        __nanosleep(64);
        if (global_ns() - t0 > timeout_ns) {
            printf("bns_p2p_wait_all: timed out waiting for flag %d to reach %llu (have %llu)\n", (int)threadIdx.x, value,
                   ld_acquire_sys(a.flag[threadIdx.x]));
            __trap();
        }
    }
}

struct ScatterAllDev {
    int32_t n_seg;
    const int32_t *inv[kMaxPeers];
    const float *recv[kMaxPeers];
    float div[kMaxPeers];
    int64_t ld_recv;
    float *G; int64_t ldg; int32_t F; int64_t n_rows;
};

// one warp per destination row: contributions of the peers are added in table order (= the reference's ring order,
// helper/feature_buffer.py:111-129), each with a true division -- bit-identical to P-1 successive scatter-adds
template <bool VEC>
__global__ void __launch_bounds__(kThreads) scatter_rows_all_kernel(ScatterAllDev a) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * kWarps;
    for (int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); row < a.n_rows; row += warps_total) {
        int32_t mine = -1;
        if (lane < a.n_seg) mine = a.inv[lane][row];
        if (__ballot_sync(0xffffffffu, mine >= 0) == 0u) continue;
        float *g = a.G + row * a.ldg;
        if (VEC) {
            for (int f = lane * 4; f < a.F; f += 128) {
                float4 v = *reinterpret_cast<float4 *>(g + f);
                for (int s = 0; s < a.n_seg; ++s) {
                    const int32_t k = __shfl_sync(0xffffffffu, mine, s);
                    if (k < 0) continue;
                    const float4 r = *reinterpret_cast<const float4 *>(a.recv[s] + (int64_t)k * a.ld_recv + f);
                    const float d = a.div[s];
                    v.x += __fdiv_rn(r.x, d); v.y += __fdiv_rn(r.y, d); v.z += __fdiv_rn(r.z, d); v.w += __fdiv_rn(r.w, d);
                }
                *reinterpret_cast<float4 *>(g + f) = v;
            }
        } else {
            for (int f0 = 0; f0 < a.F; f0 += 32) {
                const int f = f0 + lane;
                float v = f < a.F ? g[f] : 0.f;
                for (int s = 0; s < a.n_seg; ++s) {
                    const int32_t k = __shfl_sync(0xffffffffu, mine, s);
                    if (k < 0 || f >= a.F) continue;
                    v += __fdiv_rn(a.recv[s][(int64_t)k * a.ld_recv + f], a.div[s]);
                }
                if (f < a.F) g[f] = v;
            }
        }
    }
}

}  // namespace

namespace {
// see bns_p2p_create: a kernel's first launch loads its code, which synchronises the context -- never while a flag wait spins
void preload_exchange_kernels() {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, p2p_put_all_kernel<true>);
    cudaFuncGetAttributes(&fa, p2p_put_all_kernel<false>);
    cudaFuncGetAttributes(&fa, p2p_put_ids_kernel);
    cudaFuncGetAttributes(&fa, p2p_wait_all_kernel);
    cudaFuncGetAttributes(&fa, scatter_rows_all_kernel<true>);
    cudaFuncGetAttributes(&fa, scatter_rows_all_kernel<false>);
    cudaFuncGetAttributes(&fa, epoch_maps_kernel);
}
}  // namespace

extern "C" int bns_p2p_put_all_f32(bns_p2p_t *p, const bns_put_all *segs, int64_t ld_remote, const float *H, int64_t ldh,
                                   int64_t F, const int64_t *idx_cat, int32_t flag_index, int32_t ticket_index,
                                   uint64_t flag_value, const uint64_t *flag_value_dev, void *stream) {
    BNS_REQUIRE(p && segs, "bns_p2p_put_all_f32: NULL argument");
    BNS_REQUIRE(segs->n_seg >= 0 && segs->n_seg <= kMaxPeers, "bns_p2p_put_all_f32: too many segments");
    BNS_REQUIRE(flag_index >= 0 && flag_index < p->n_flags, "bns_p2p_put_all_f32: bad flag index");
    BNS_REQUIRE(ticket_index >= 0 && ticket_index < p->world + 16, "bns_p2p_put_all_f32: bad ticket index");
    BNS_REQUIRE(F > 0 && ldh >= F && ld_remote >= F, "bns_p2p_put_all_f32: bad shape");
    if (segs->n_seg == 0) return BNS_OK;
    PutAllDev a;
    a.n_seg = segs->n_seg;
    bool vec = F % 4 == 0 && ldh % 4 == 0 && ld_remote % 4 == 0 && (reinterpret_cast<uintptr_t>(H) & 15u) == 0;
    for (int s = 0; s <= segs->n_seg; ++s) a.row_begin[s] = segs->row_begin[s];
    for (int s = 0; s < segs->n_seg; ++s) {
        const int peer = segs->peer[s];
        const int64_t k = segs->row_begin[s + 1] - segs->row_begin[s];
        BNS_REQUIRE(peer >= 0 && peer < p->world && peer != p->rank, "bns_p2p_put_all_f32: bad peer %d", peer);
        BNS_REQUIRE(p->peer_slab[peer] && p->peer_flags[peer], "bns_p2p_put_all_f32: peer %d not connected", peer);
        BNS_REQUIRE(k >= 0, "bns_p2p_put_all_f32: negative row count");
        BNS_REQUIRE(k == 0 || segs->div[s] != 0.f, "bns_p2p_put_all_f32: division by zero");
        BNS_REQUIRE(segs->remote_off[s] % 16 == 0 &&
                        segs->remote_off[s] + (size_t)k * ld_remote * 4 <= p->peer_slab_bytes[peer],
                    "bns_p2p_put_all_f32: remote range of segment %d outside peer %d's slab", s, peer);
        a.remote[s] = reinterpret_cast<float *>(p->peer_slab[peer] + segs->remote_off[s]);
        a.flag[s] = p->peer_flags[peer] + flag_index;
        a.src_begin[s] = segs->src_begin[s];
        a.div[s] = k == 0 ? 1.f : segs->div[s];
    }
    const int64_t total = segs->row_begin[segs->n_seg];
    BNS_REQUIRE(total == 0 || H, "bns_p2p_put_all_f32: NULL source");
    a.H = H; a.ldh = ldh; a.F = (int32_t)F; a.idx = idx_cat; a.ld_remote = ld_remote;
    a.flag_value = flag_value; a.flag_value_dev = reinterpret_cast<const unsigned long long *>(flag_value_dev);
    a.ticket = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(p->flags) + align256((size_t)p->n_flags * 8)) + ticket_index;
    const unsigned grid = rows_grid(total);
    if (vec) p2p_put_all_kernel<true><<<grid, kThreads, 0, as_stream(stream)>>>(a);
    else p2p_put_all_kernel<false><<<grid, kThreads, 0, as_stream(stream)>>>(a);
    ++g_launches;
    BNS_CUDA(cudaGetLastError());
    return BNS_OK;
}

extern "C" int bns_p2p_put_ids_i64(bns_p2p_t *p, int32_t n_seg, const int64_t *begin, const int32_t *peers,
                                   const uint64_t *remote_off, const int64_t *ids_cat, int32_t flag_index,
                                   int32_t ticket_index, uint64_t flag_value, const uint64_t *flag_value_dev, void *stream) {
    BNS_REQUIRE(p && begin && peers && remote_off, "bns_p2p_put_ids_i64: NULL argument");
    BNS_REQUIRE(n_seg >= 0 && n_seg <= kMaxPeers, "bns_p2p_put_ids_i64: too many segments");
    BNS_REQUIRE(flag_index >= 0 && flag_index < p->n_flags, "bns_p2p_put_ids_i64: bad flag index");
    BNS_REQUIRE(ticket_index >= 0 && ticket_index < p->world + 16, "bns_p2p_put_ids_i64: bad ticket index");
    if (n_seg == 0) return BNS_OK;
    PutIdsDev a;
    a.n_seg = n_seg;
    for (int s = 0; s <= n_seg; ++s) a.begin[s] = begin[s];
    for (int s = 0; s < n_seg; ++s) {
        const int peer = peers[s];
        const int64_t k = begin[s + 1] - begin[s];
        BNS_REQUIRE(peer >= 0 && peer < p->world && peer != p->rank, "bns_p2p_put_ids_i64: bad peer %d", peer);
        BNS_REQUIRE(p->peer_slab[peer] && p->peer_flags[peer], "bns_p2p_put_ids_i64: peer %d not connected", peer);
        BNS_REQUIRE(k >= 0 && remote_off[s] % 8 == 0 && remote_off[s] + (size_t)k * 8 <= p->peer_slab_bytes[peer],
                    "bns_p2p_put_ids_i64: remote range of segment %d outside peer %d's slab", s, peer);
        a.remote[s] = reinterpret_cast<int64_t *>(p->peer_slab[peer] + remote_off[s]);
        a.flag[s] = p->peer_flags[peer] + flag_index;
    }
    BNS_REQUIRE(begin[n_seg] == 0 || ids_cat, "bns_p2p_put_ids_i64: NULL source");
    a.src = ids_cat;
    a.flag_value = flag_value; a.flag_value_dev = reinterpret_cast<const unsigned long long *>(flag_value_dev);
    a.ticket = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(p->flags) + align256((size_t)p->n_flags * 8)) + ticket_index;
    const int64_t total = begin[n_seg];
    int64_t want = (total + 255) / 256;
    if (want < 1) want = 1;
    if (want > 64) want = 64;
    p2p_put_ids_kernel<<<(unsigned)want, 256, 0, as_stream(stream)>>>(a);
    ++g_launches;
    BNS_CUDA(cudaGetLastError());
    return BNS_OK;
}

extern "C" int bns_p2p_wait_all(bns_p2p_t *p, int32_t n, const int32_t *flag_indices, uint64_t flag_value,
                                const uint64_t *flag_value_dev, void *stream) {
    BNS_REQUIRE(p && (n == 0 || flag_indices), "bns_p2p_wait_all: NULL argument");
    BNS_REQUIRE(n >= 0 && n <= kMaxPeers, "bns_p2p_wait_all: too many flags");
    if (n == 0) return BNS_OK;
    WaitAllDev a;
    a.n = n;
    for (int i = 0; i < n; ++i) {
        BNS_REQUIRE(flag_indices[i] >= 0 && flag_indices[i] < p->n_flags, "bns_p2p_wait_all: bad flag index");
        a.flag[i] = p->flags + flag_indices[i];
    }
    p2p_wait_all_kernel<<<1, 32, 0, as_stream(stream)>>>(a, flag_value, reinterpret_cast<const unsigned long long *>(flag_value_dev),
                                                         20ull * 1000000000ull);
    ++g_launches;
    BNS_CUDA(cudaGetLastError());
    return BNS_OK;
}

// G[r] += sum over segments s (in order) of recv_s[inv_s[r]] / div_s, for the rows r some peer selected
extern "C" int bns_scatter_rows_all_f32(float *G, int64_t ldg, int64_t n_rows, int64_t F, int32_t n_seg,
                                        const int32_t *const *inv, const float *const *recv, int64_t ld_recv,
                                        const float *div, void *stream) {
    BNS_REQUIRE(n_seg >= 0 && n_seg <= kMaxPeers, "bns_scatter_rows_all_f32: too many segments");
    if (n_seg == 0 || n_rows == 0) return BNS_OK;
    BNS_REQUIRE(G && inv && recv && div && F > 0 && ldg >= F && ld_recv >= F, "bns_scatter_rows_all_f32: bad argument");
    ScatterAllDev a;
    a.n_seg = n_seg;
    bool vec = F % 4 == 0 && ldg % 4 == 0 && ld_recv % 4 == 0 && (reinterpret_cast<uintptr_t>(G) & 15u) == 0;
    for (int s = 0; s < n_seg; ++s) {
        BNS_REQUIRE(inv[s] && recv[s] && div[s] != 0.f, "bns_scatter_rows_all_f32: bad segment %d", s);
        a.inv[s] = inv[s]; a.recv[s] = recv[s]; a.div[s] = div[s];
        vec = vec && (reinterpret_cast<uintptr_t>(recv[s]) & 15u) == 0;
    }
    a.ld_recv = ld_recv; a.G = G; a.ldg = ldg; a.F = (int32_t)F; a.n_rows = n_rows;
    if (vec) scatter_rows_all_kernel<true><<<rows_grid(n_rows), kThreads, 0, as_stream(stream)>>>(a);
    else scatter_rows_all_kernel<false><<<rows_grid(n_rows), kThreads, 0, as_stream(stream)>>>(a);
    ++g_launches;
    BNS_CUDA(cudaGetLastError());
    return BNS_OK;
}

// =====================================================================================================================
// per-epoch compaction of a column-mapped matrix (the halo matrix A_out) to the sampled columns
// =====================================================================================================================
namespace {

// One warp per static chunk (<= chunk_nnz entries of one row): the live entries (col_map >= 0) are written, already
// mapped to rows of X, to cidx[chunk_start .. chunk_start + cnt) -- IN PLACE of the chunk's own index range, so there is
// no scan -- and cnt to chunk_cnt.  The order of the live entries is the CSR order: the SpMM that walks the compacted
// chunks adds exactly the same numbers in exactly the same order as the col_map kernel.
__global__ void __launch_bounds__(kThreads) compact_cols_kernel(const int64_t *__restrict__ indptr,
                                                                const int32_t *__restrict__ indices,
                                                                const int32_t *__restrict__ chunk_row,
                                                                const int64_t *__restrict__ chunk_start, int64_t n_chunks,
                                                                int32_t chunk_nnz, const int32_t *__restrict__ col_map,
                                                                int32_t n_direct, const float *__restrict__ col_scale,
                                                                int32_t *__restrict__ cidx, float *__restrict__ cw,
                                                                int32_t *__restrict__ cpos, int32_t *__restrict__ chunk_cnt) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * kWarps;
    for (int64_t c = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); c < n_chunks; c += warps_total) {
        const int32_t row = chunk_row[c];
        const int64_t s = chunk_start[c];
        int64_t e = indptr[row + 1];
        if (e > s + chunk_nnz) e = s + chunk_nnz;
        int32_t off = 0;
        constexpr int U = 4;              // 4 x 32 ids and their 4 x 32 slot look-ups in flight per warp
        for (int64_t k0 = s; k0 < e; k0 += 32 * U) {
            int32_t orig[U], col[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t k = k0 + 32 * u + lane;
                orig[u] = k < e ? ld_stream_i32(indices + k) : -1;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                col[u] = -1;
                if (orig[u] >= 0) col[u] = orig[u] >= n_direct ? __ldg(col_map + (orig[u] - n_direct)) : orig[u];
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned m = __ballot_sync(0xffffffffu, col[u] >= 0);
                if (col[u] >= 0) {
                    const int pos = off + __popc(m & ((1u << lane) - 1u));
                    cidx[s + pos] = col[u];
                    if (cw) cw[s + pos] = __ldg(col_scale + orig[u]);
                    if (cpos) cpos[s + pos] = (int32_t)(k0 + 32 * u + lane);       // where the entry sits in the CSR
                }
                off += __popc(m);
            }
        }
        if (lane == 0) chunk_cnt[c] = off;
    }
}

}  // namespace

extern "C" int bns_graph_compact_cols(const bns_graph_t *g, const int32_t *col_map, int64_t n_direct, const float *col_scale,
                                      int32_t *cidx /*[nnz]*/, float *cw /*[nnz] or NULL*/, int32_t *cpos /*[nnz] or NULL*/,
                                      int32_t *chunk_cnt /*[n_chunks]*/, void *stream) {
    BNS_REQUIRE(!cpos || (g && g->nnz < INT32_MAX), "bns_graph_compact_cols: cpos needs nnz < 2^31");
    BNS_REQUIRE(g && col_map && cidx && chunk_cnt, "bns_graph_compact_cols: NULL argument");
    BNS_REQUIRE((cw == nullptr) == (col_scale == nullptr), "bns_graph_compact_cols: cw and col_scale go together");
    BNS_REQUIRE(n_direct >= 0 && n_direct <= g->n_cols, "bns_graph_compact_cols: n_direct out of range");
    if (g->n_chunks == 0) return BNS_OK;
    int64_t want = (g->n_chunks + kWarps - 1) / kWarps, cap = (int64_t)sm_count() * 8;
    compact_cols_kernel<<<(unsigned)(want < cap ? want : cap), kThreads, 0, as_stream(stream)>>>(
        g->indptr, g->indices, g->chunk_row, g->chunk_start, g->n_chunks, g->chunk_nnz, col_map, (int32_t)n_direct, col_scale,
        cidx, cw, cpos, chunk_cnt);
    ++g_launches;
    BNS_CUDA(cudaGetLastError());
    return BNS_OK;
}

// =====================================================================================================================
// SyncBatchNorm (--norm batch, module/sync_bn.py:7-56 of the reference): batch statistics over ALL partitions.
// Per step and layer: ONE pass producing both moments (sum x, sum x^2) per column, one packed [2F] all-reduce (done by
// the caller), ONE normalise+affine pass; mirrored in backward (sum dy, sum dy*x_hat -> packed all-reduce -> dx).  The
// reference issues four [F] all-reduces and ~12 element-wise ATen launches for the same.
// =====================================================================================================================
namespace {

// MODE 0: (x, x^2)      MODE 1: (dy, dy * x_hat) with x_hat = (x - mean) * rstd
// block b sums a contiguous row range; thread (rg, c) owns float4 column c of every RG-th row of it; fixed order.
template <int MODE>
__global__ void __launch_bounds__(kThreads) colsum2_partial_kernel(const float *__restrict__ A, int64_t lda,
                                                                  const float *__restrict__ X, int64_t ldx, int64_t rows, int CV,
                                                                  const float4 *__restrict__ mean, const float4 *__restrict__ rstd,
                                                                  float4 *__restrict__ partial) {
    __shared__ float4 s0[kThreads], s1[kThreads];
    const int RG = kThreads / CV;
    const int rg = threadIdx.x / CV, c = threadIdx.x % CV;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    if (rg < RG) {
        float4 mu = a0, rs = a0;
        if (MODE == 1) { mu = mean[c]; rs = rstd[c]; }
        const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
        const int64_t r0 = (int64_t)blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
        for (int64_t r = r0 + rg; r < r1; r += RG) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(A + r * lda) + c);
            a0.x += v.x; a0.y += v.y; a0.z += v.z; a0.w += v.w;
            if (MODE == 0) {
                a1.x = fmaf(v.x, v.x, a1.x); a1.y = fmaf(v.y, v.y, a1.y); a1.z = fmaf(v.z, v.z, a1.z); a1.w = fmaf(v.w, v.w, a1.w);
            } else {
                const float4 x = __ldg(reinterpret_cast<const float4 *>(X + r * ldx) + c);
                a1.x = fmaf(v.x, (x.x - mu.x) * rs.x, a1.x); a1.y = fmaf(v.y, (x.y - mu.y) * rs.y, a1.y);
                a1.z = fmaf(v.z, (x.z - mu.z) * rs.z, a1.z); a1.w = fmaf(v.w, (x.w - mu.w) * rs.w, a1.w);
            }
        }
    }
    s0[threadIdx.x] = a0; s1[threadIdx.x] = a1;
    __syncthreads();
    if (rg == 0) {
        for (int g = 1; g < RG; ++g) {
            const float4 u = s0[g * CV + c], w = s1[g * CV + c];
            a0.x += u.x; a0.y += u.y; a0.z += u.z; a0.w += u.w;
            a1.x += w.x; a1.y += w.y; a1.z += w.z; a1.w += w.w;
        }
        partial[(int64_t)blockIdx.x * 2 * CV + c] = a0;
        partial[(int64_t)blockIdx.x * 2 * CV + CV + c] = a1;
    }
}

// y = (x - mean) * rstd * w + b with mean = S1 / n, var = (S2 - mean * S1) / n (the reference's one-pass variance,
// sync_bn.py:19-20); block 0 also stores mean / rstd for the backward and moves the running statistics
__global__ void __launch_bounds__(kThreads) bn_apply_kernel(const float *__restrict__ x, int64_t ldx, int64_t rows, int32_t F,
                                                           const float *__restrict__ sums, float n, float eps,
                                                           const float *__restrict__ w, const float *__restrict__ b,
                                                           float momentum, float *running_mean, float *running_var,
                                                           float *__restrict__ y, int64_t ldy, float *mean_out, float *rstd_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * kWarps;
    if (blockIdx.x == 0) {
        for (int f = threadIdx.x; f < F; f += kThreads) {
            const float mu = sums[f] / n, var = (sums[F + f] - mu * sums[f]) / n;
            mean_out[f] = mu;
            rstd_out[f] = 1.f / sqrtf(var + eps);
            if (running_mean) {
                running_mean[f] = running_mean[f] * (1.f - momentum) + mu * momentum;
                running_var[f] = running_var[f] * (1.f - momentum) + var * momentum;
            }
        }
    }
    for (int f0 = lane * 4; f0 < F; f0 += 128) {
        float mu[4], rs[4], ww[4], bb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = f0 + i;
            mu[i] = sums[f] / n;
            const float var = (sums[F + f] - mu[i] * sums[f]) / n;
            rs[i] = 1.f / sqrtf(var + eps);
            ww[i] = w[f]; bb[i] = b[f];
        }
        for (int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); row < rows; row += warps_total) {
            const float4 v = *reinterpret_cast<const float4 *>(x + row * ldx + f0);
            float4 o;
            o.x = fmaf((v.x - mu[0]) * rs[0], ww[0], bb[0]); o.y = fmaf((v.y - mu[1]) * rs[1], ww[1], bb[1]);
            o.z = fmaf((v.z - mu[2]) * rs[2], ww[2], bb[2]); o.w = fmaf((v.w - mu[3]) * rs[3], ww[3], bb[3]);
            *reinterpret_cast<float4 *>(y + row * ldy + f0) = o;
        }
    }
}

// dx = (w / n) * rstd * (n * dy - dbias - x_hat * dweight)      (sync_bn.py:51-54)
__global__ void __launch_bounds__(kThreads) bn_bwd_kernel(const float *__restrict__ dy, int64_t lddy, const float *__restrict__ x,
                                                         int64_t ldx, int64_t rows, int32_t F, const float *__restrict__ mean,
                                                         const float *__restrict__ rstd, const float *__restrict__ w,
                                                         const float *__restrict__ sums, float n, float *__restrict__ dx,
                                                         int64_t lddx) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * kWarps;
    for (int f0 = lane * 4; f0 < F; f0 += 128) {
        float mu[4], rs[4], k[4], db[4], dw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = f0 + i;
            mu[i] = mean[f]; rs[i] = rstd[f]; k[i] = (w[f] / n) * rs[i]; db[i] = sums[f]; dw[i] = sums[F + f];
        }
        for (int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); row < rows; row += warps_total) {
            const float4 g = *reinterpret_cast<const float4 *>(dy + row * lddy + f0);
            const float4 v = *reinterpret_cast<const float4 *>(x + row * ldx + f0);
            float4 o;
            o.x = k[0] * (n * g.x - db[0] - (v.x - mu[0]) * rs[0] * dw[0]);
            o.y = k[1] * (n * g.y - db[1] - (v.y - mu[1]) * rs[1] * dw[1]);
            o.z = k[2] * (n * g.z - db[2] - (v.z - mu[2]) * rs[2] * dw[2]);
            o.w = k[3] * (n * g.w - db[3] - (v.w - mu[3]) * rs[3] * dw[3]);
            *reinterpret_cast<float4 *>(dx + row * lddx + f0) = o;
        }
    }
}

inline bool bn_shape_ok(const void *a, int64_t ld, int64_t F) {
    return F > 0 && F % 4 == 0 && F <= kColsumMaxCols && ld % 4 == 0 && ld >= F && (reinterpret_cast<uintptr_t>(a) & 15u) == 0;
}

}  // namespace

extern "C" size_t bns_bn_workspace_bytes(int64_t F) { return 2 * bns_colsum_workspace_bytes(F); }

// mode 0: out[0:F] = column sums of A, out[F:2F] = column sums of A^2            (forward moments; X, mean, rstd unused)
// mode 1: out[0:F] = column sums of A (= dy), out[F:2F] = column sums of A * x_hat, x_hat = (X - mean) * rstd
// rows == 0 writes zeros.  Deterministic (fixed two-pass order).  ws: bns_bn_workspace_bytes(F).
extern "C" int bns_bn_colsums_f32(int mode, const float *A, int64_t lda, const float *X, int64_t ldx, int64_t rows, int64_t F,
                                  const float *mean, const float *rstd, float *out /*[2F]*/, void *ws, size_t ws_bytes,
                                  void *stream) {
    BNS_REQUIRE(mode == 0 || mode == 1, "bns_bn_colsums_f32: mode must be 0 or 1");
    BNS_REQUIRE(out && (reinterpret_cast<uintptr_t>(out) & 15u) == 0, "bns_bn_colsums_f32: bad output");
    BNS_REQUIRE(rows >= 0 && F > 0 && F % 4 == 0 && F <= kColsumMaxCols, "bns_bn_colsums_f32: need F %% 4 == 0, F <= 1024");
    cudaStream_t st = as_stream(stream);
    if (rows == 0) {
        BNS_CUDA(cudaMemsetAsync(out, 0, 2 * F * sizeof(float), st));
        return BNS_OK;
    }
    BNS_REQUIRE(A && bn_shape_ok(A, lda, F), "bns_bn_colsums_f32: A must have 16-byte aligned rows");
    if (mode == 1)
        BNS_REQUIRE(X && mean && rstd && bn_shape_ok(X, ldx, F) && ((reinterpret_cast<uintptr_t>(mean) | reinterpret_cast<uintptr_t>(rstd)) & 15u) == 0,
                    "bns_bn_colsums_f32: mode 1 needs X, mean, rstd (16-byte aligned)");
    if (!ws || ws_bytes < bns_bn_workspace_bytes(F) || (reinterpret_cast<uintptr_t>(ws) & 15u))
        return fail(BNS_E_WORKSPACE, "bns_bn_colsums_f32: workspace %zu bytes < %zu needed", ws_bytes, bns_bn_workspace_bytes(F));
    const int CV = (int)(F / 4);
    int blocks = colsum_blocks();
    if ((int64_t)blocks > rows) blocks = (int)rows;
    if (mode == 0)
        colsum2_partial_kernel<0><<<blocks, kThreads, 0, st>>>(A, lda, nullptr, 0, rows, CV, nullptr, nullptr, reinterpret_cast<float4 *>(ws));
    else
        colsum2_partial_kernel<1><<<blocks, kThreads, 0, st>>>(A, lda, X, ldx, rows, CV, reinterpret_cast<const float4 *>(mean),
                                                              reinterpret_cast<const float4 *>(rstd), reinterpret_cast<float4 *>(ws));
    colsum_final_kernel<<<(2 * CV + kWarps - 1) / kWarps, kThreads, 0, st>>>(reinterpret_cast<const float4 *>(ws), blocks, 2 * CV,
                                                                             reinterpret_cast<float4 *>(out), nullptr);
    g_launches += 2;
    BNS_CUDA(cudaGetLastError());
    return BNS_OK;
}

extern "C" int bns_bn_apply_f32(const float *x, int64_t ldx, int64_t rows, int64_t F, const float *sums /*[2F], all ranks*/,
                                float whole_size, float eps, const float *weight, const float *bias, float momentum,
                                float *running_mean, float *running_var, float *y, int64_t ldy, float *mean_out, float *rstd_out,
                                void *stream) {
    BNS_REQUIRE(rows >= 0 && F > 0 && F % 4 == 0 && F <= kColsumMaxCols, "bns_bn_apply_f32: need F %% 4 == 0, F <= 1024");
    BNS_REQUIRE(sums && weight && bias && mean_out && rstd_out && whole_size > 0.f, "bns_bn_apply_f32: NULL argument");
    BNS_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "bns_bn_apply_f32: running stats go together");
    BNS_REQUIRE(rows == 0 || (x && y && bn_shape_ok(x, ldx, F) && bn_shape_ok(y, ldy, F)), "bns_bn_apply_f32: bad matrix");
    bn_apply_kernel<<<ln_grid(rows > 0 ? rows : 1), kThreads, 0, as_stream(stream)>>>(x, ldx, rows, (int32_t)F, sums, whole_size, eps,
                                                                                    weight, bias, momentum, running_mean,
                                                                                    running_var, y, ldy, mean_out, rstd_out);
    ++g_launches;
    BNS_CUDA(cudaGetLastError());
    return BNS_OK;
}

extern "C" int bns_bn_bwd_f32(const float *dy, int64_t lddy, const float *x, int64_t ldx, int64_t rows, int64_t F,
                              const float *mean, const float *rstd, const float *weight,
                              const float *sums /*[2F]: sum dy, sum dy * x_hat over all ranks*/, float whole_size, float *dx,
                              int64_t lddx, void *stream) {
    BNS_REQUIRE(rows >= 0 && F > 0 && F % 4 == 0 && F <= kColsumMaxCols, "bns_bn_bwd_f32: need F %% 4 == 0, F <= 1024");
    if (rows == 0) return BNS_OK;
    BNS_REQUIRE(dy && x && dx && mean && rstd && weight && sums && whole_size > 0.f, "bns_bn_bwd_f32: NULL argument");
    BNS_REQUIRE(bn_shape_ok(dy, lddy, F) && bn_shape_ok(x, ldx, F) && bn_shape_ok(dx, lddx, F), "bns_bn_bwd_f32: bad matrix");
    bn_bwd_kernel<<<ln_grid(rows), kThreads, 0, as_stream(stream)>>>(dy, lddy, x, ldx, rows, (int32_t)F, mean, rstd, weight, sums,
                                                                     whole_size, dx, lddx);
    ++g_launches;
    BNS_CUDA(cudaGetLastError());
    return BNS_OK;
}
