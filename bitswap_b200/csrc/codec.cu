// The bits-back recursions, device resident (sm_100a).
//
// Reference being replaced (fhkingma/bitswap @ dfe0bf7d):
//   Bit-Swap sender   cifar_compress.py:175-204, 244-250      receiver  :283-317
//   BB-ANS  sender    cifar_compress.py:205-242, 244-250      receiver  :319-352
// One call codes ONE image per stream for `count` streams.  Every step is a kernel enqueued on the
// caller's stream; symbols, centres, mu/sigma and the ANS state stay in HBM between levels -- the
// reference's 2*nz GPU<->CPU round trips per image (cifar_compress.py:42-43,66) do not exist here.
#include <vector>
#include "bsw_common.cuh"
#include "nets.cuh"

struct bsw_codec {
    bsw_model *m;
    bsw_bins *b;
    int max_batch, nz, zdim, xdim, q, S;
    float *given, *mu, *scale;          // [max_batch, max(xdim, zdim)]
    int16_t *sym[2];                    // ping-pong latent symbols [max_batch, zdim]
    int16_t *xsym;                      // [max_batch, xdim] pixels as int16 symbols
    std::vector<int16_t *> zs;          // BB-ANS: all nz latents
    void *scratch;                      // two-phase coder: pairs / coarse cdf (/ full integer table) of one level
    size_t scratch_bytes;
    int two_phase;                      // 1: ans_rows.cu path (default), 0: fused one-warp-per-stream kernels
    uint32_t *priorP, *priorC;
    uint32_t *priorCoarse = nullptr;    // prior cdf at every 32nd bin + dummy fix: lets the prior pop use k_pop_full
    uint2 *priorFix = nullptr;          // Logistic(0,1) prior tables over zendpoints[-1], shared by all streams
    int64_t launches;
    BswProf prof;
    // overlap: convs go to a high-priority internal stream, coder kernels to a low-priority one, chained by events, so
    // that (with several codecs in flight) tensor-bound conv CTAs and FP64-bound coder CTAs share SMs
    cudaStream_t st_hi = nullptr, st_lo = nullptr;
    cudaEvent_t ev[8];
    int ev_next = 0;
    int dual_stream = 0;     // measured on B200: no gain (lanes=4: 350 ms/step without, 378 with) -- kept as an option
};

// internal int16-symbol variants of the table-driven coder (ans_kernels.cu)
int bsw_ans_push_i16(bsw_streams *s, int first, int count, const uint32_t *P, const uint32_t *C, int64_t pss, int64_t css,
                     const int16_t *sym, int64_t L, int S, int bits, cudaStream_t st);
int bsw_ans_pop_i16(bsw_streams *s, int first, int count, const uint32_t *P, const uint32_t *C, int64_t pss, int64_t css,
                    int16_t *sym, int64_t L, int S, int bits, cudaStream_t st);

size_t bsw_rows_scratch_bytes(int count, int64_t L);
int bsw_streams_mark_rest(bsw_streams *s, int first, int count, cudaStream_t st);
int bsw_rows_mode();
int bsw_prior_coarse(const uint32_t *C, int64_t L, int S, uint32_t *coarse, uint2 *fix, cudaStream_t st);
int bsw_pop_shared_table(bsw_streams *s, int first, int count, const uint32_t *P, const uint32_t *coarse, const uint2 *fix,
                         int16_t *sym, int64_t L, int S, int bits, cudaStream_t st);
int bsw_logistic_2p(int phase, bool pop, bsw_streams *s, int first, int count, const float *mu, int64_t mss, const float *sc,
                    int64_t sss, const double *endp, int64_t ers, const void *meta, int16_t *sym, int64_t L, int S, int bits, int q,
                    void *scratch, size_t scratch_bytes, cudaStream_t st);

__global__ void k_u8_to_i16(const uint8_t *__restrict__ in, int16_t *__restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
__global__ void k_i16_to_u8(const int16_t *__restrict__ in, uint8_t *__restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint8_t)in[i];
}

extern "C" int bsw_codec_create(bsw_codec **out, bsw_model *m, bsw_bins *b, int max_batch) {
    BSW_REQUIRE(out && m && b && max_batch >= 1, "bsw_codec_create: bad arguments");
    BSW_REQUIRE(m->finalized, "bsw_codec_create: model not finalized");
    BSW_REQUIRE(b->nz == m->d.nz && b->zdim == m->zdim, "bsw_codec_create: bins do not match the model");
    BSW_REQUIRE(max_batch <= m->d.max_batch, "bsw_codec_create: max_batch exceeds the model's");
    bsw_codec *c = new bsw_codec();
    c->m = m; c->b = b; c->max_batch = max_batch;
    c->nz = b->nz; c->zdim = b->zdim; c->xdim = m->xdim; c->q = b->q; c->S = b->S;
    size_t dim = (size_t)(c->xdim > c->zdim ? c->xdim : c->zdim);
    BSW_CUDA(cudaMalloc(&c->given, sizeof(float) * dim * max_batch));
    BSW_CUDA(cudaMalloc(&c->mu, sizeof(float) * dim * max_batch));
    BSW_CUDA(cudaMalloc(&c->scale, sizeof(float) * dim * max_batch));
    BSW_CUDA(cudaMalloc(&c->sym[0], sizeof(int16_t) * c->zdim * max_batch));
    BSW_CUDA(cudaMalloc(&c->sym[1], sizeof(int16_t) * c->zdim * max_batch));
    BSW_CUDA(cudaMalloc(&c->xsym, sizeof(int16_t) * c->xdim * max_batch));
    c->zs.assign(c->nz, nullptr);
    {
        size_t a = bsw_rows_scratch_bytes(max_batch, c->zdim), b2 = bsw_rows_scratch_bytes(max_batch, c->xdim);
        c->scratch_bytes = a > b2 ? a : b2;
        BSW_CUDA(cudaMalloc(&c->scratch, c->scratch_bytes));
        c->two_phase = 1;
    }
    // prior tables: Logistic(0,1) over the top level's endpoints, identical for every stream and image
    // (cifar_compress.py:245-247) -> built once, with the same float64 kernel math as every other table.
    BSW_CUDA(cudaMalloc(&c->priorP, sizeof(uint32_t) * (size_t)c->zdim * c->S));
    BSW_CUDA(cudaMalloc(&c->priorC, sizeof(uint32_t) * (size_t)c->zdim * (c->S + 1)));
    double ms[2] = {0.0, 1.0}, *dms = nullptr;
    BSW_CUDA(cudaMalloc(&dms, sizeof(ms)));
    BSW_CUDA(cudaMemcpy(dms, ms, sizeof(ms), cudaMemcpyHostToDevice));
    const double *zend = b->zend + (size_t)(c->nz - 1) * c->zdim * c->S;
    int rc = bsw_logistic_tables(zend, c->S, dms, dms + 1, 0, c->zdim, c->S, 31, c->q, c->priorP, c->priorC, nullptr);
    if (!rc && c->S >= 128) {
        BSW_CUDA(cudaMalloc(&c->priorCoarse, sizeof(uint32_t) * (size_t)c->zdim * (c->S / 32)));
        BSW_CUDA(cudaMalloc(&c->priorFix, sizeof(uint2) * (size_t)c->zdim));
        rc = bsw_prior_coarse(c->priorC, c->zdim, c->S, c->priorCoarse, c->priorFix, nullptr);
    }
    BSW_CUDA(cudaDeviceSynchronize());
    cudaFree(dms);
    if (rc) return rc;
    c->launches = 0;
    int lo_p = 0, hi_p = 0;
    BSW_CUDA(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));          // (numerically lower = higher priority)
    BSW_CUDA(cudaStreamCreateWithPriority(&c->st_hi, cudaStreamNonBlocking, hi_p));
    BSW_CUDA(cudaStreamCreateWithPriority(&c->st_lo, cudaStreamNonBlocking, lo_p));
    for (auto &e : c->ev) BSW_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    *out = c;
    return BSW_OK;
}

extern "C" int bsw_codec_destroy(bsw_codec *c) {
    if (!c) return BSW_OK;
    cudaFree(c->given); cudaFree(c->mu); cudaFree(c->scale);
    cudaFree(c->sym[0]); cudaFree(c->sym[1]); cudaFree(c->xsym);
    for (auto p : c->zs) cudaFree(p);
    cudaFree(c->priorP); cudaFree(c->priorC); cudaFree(c->scratch); cudaFree(c->priorCoarse); cudaFree(c->priorFix);
    if (c->st_hi) cudaStreamDestroy(c->st_hi);
    if (c->st_lo) cudaStreamDestroy(c->st_lo);
    for (auto &e : c->ev) cudaEventDestroy(e);
    delete c;
    return BSW_OK;
}
extern "C" int64_t bsw_codec_last_launches(const bsw_codec *c) { return c ? c->launches : 0; }

// Per-category kernel timing for bench.py's roofline (CUDA events on the launching stream).
/* 1 (default): two-phase coder (parallel row tables + serial coder); 0: fused one-warp-per-stream kernels. */
/* 1: convs on a high-priority internal stream, coder kernels on a low-priority one (event-chained);
 * 0 (default): everything on the caller's stream. */
extern "C" int bsw_codec_set_dual_stream(bsw_codec *c, int on) {
    BSW_REQUIRE(c, "null codec");
    c->dual_stream = on < 0 ? 0 : (on > 2 ? 2 : on);
    return BSW_OK;
}

extern "C" int bsw_codec_set_two_phase(bsw_codec *c, int on) {
    BSW_REQUIRE(c, "null codec");
    c->two_phase = on ? 1 : 0;
    return BSW_OK;
}

extern "C" int bsw_codec_profile(bsw_codec *c, int enable, double *ms_out, int64_t *n_out) {
    BSW_REQUIRE(c, "null codec");
    c->prof.collect();
    if (ms_out) for (int i = 0; i < CAT_COUNT; ++i) ms_out[i] = c->prof.ms[i];
    if (n_out) for (int i = 0; i < CAT_COUNT; ++i) n_out[i] = c->prof.n[i];
    if (enable >= 0) { c->prof.reset(); c->prof.on = enable != 0; }
    return BSW_OK;
}

#define RC(call) do { if (int rc_ = (call)) return rc_; } while (0)

namespace {
struct Ctx {
    bsw_codec *c; bsw_streams *s; int first, count; cudaStream_t st;      // st = stream the next kernel goes to
    int nl = 0;
    cudaStream_t caller = nullptr;
    // move the chain to `target`: everything enqueued so far (on st) happens-before what follows on target
    void use(cudaStream_t target) {
        if (target == st) return;
        cudaEvent_t e = c->ev[c->ev_next];
        c->ev_next = (c->ev_next + 1) & 7;
        cudaEventRecord(e, st);
        cudaStreamWaitEvent(target, e, 0);
        st = target;
    }
    // dual_stream 1: convs on the high-priority internal stream, coder kernels on the low-priority one (measured: slower).
    // dual_stream 2: only the serial phase-B coder kernels (one warp per stream, latency-bound, a few hundred small CTAs) move
    //   to the high-priority stream, so that in a multi-lane run their CTAs are dispatched ahead of the queued throughput
    //   CTAs of the other lanes instead of waiting behind a k_rows grid that owns every register of every SM.
    void begin(cudaStream_t user) { caller = user; st = user; if (c->dual_stream == 1) use(c->st_lo); }
    void end() { use(caller); }
    void conv_stream() { if (c->dual_stream == 1) use(c->st_hi); else if (c->dual_stream == 2) use(caller); }
    void coder_stream() { if (c->dual_stream == 1) use(c->st_lo); else if (c->dual_stream == 2) use(caller); }
    void serial_stream() { if (c->dual_stream == 2) use(c->st_hi); }
    const double *zend(int lvl) const { return c->b->zend + (size_t)lvl * c->zdim * c->S; }
    // row metadata of a level whose rows are all uniform grids (-> affine-row kernels), else NULL (-> generic kernels)
    const void *zmeta(int lvl) const {
        return (c->b->zaffine[lvl] && bsw_rows_mode() != 0) ? (const uint8_t *)c->b->zmeta + (size_t)32 * lvl * c->zdim : nullptr;
    }
    const void *xmeta() const { return (c->b->xaffine && bsw_rows_mode() != 0) ? c->b->xmeta : nullptr; }
    int infer(int zi) { conv_stream(); return bsw_model_run(c->m, true, zi, c->given, count, c->mu, c->scale, 1, st, &nl, &c->prof); }
    int generate(int zi) { conv_stream(); return bsw_model_run(c->m, false, zi, c->given, count, c->mu, c->scale, 0, st, &nl, &c->prof); }
    int gather_x(const uint8_t *x) {
        coder_stream();
        ++nl;
        c->prof.begin(CAT_MISC, st);
        int rc = bsw_gather_xcentres(x, c->given, (int64_t)count * c->xdim, st);
        c->prof.end(st);
        return rc;
    }
    int gather_x16(const int16_t *x);
    int gather_z(int lvl, const int16_t *sym) {
        coder_stream();
        ++nl;
        c->prof.begin(CAT_MISC, st);
        int rc = bsw_gather_zcentres(c->b, lvl, sym, c->given, count, st);
        c->prof.end(st);
        return rc;
    }
    // q(z_{zi+1} | .) / p(z_zi | .) tables over level `lvl` endpoints
    int pop_z(int lvl, int16_t *sym) {
        coder_stream();
        if (!c->two_phase) {
            ++nl;
            c->prof.begin(CAT_POP_Z, st);
            int rc = bsw_logistic_pop(s, first, count, c->mu, c->zdim, c->scale, c->zdim, zend(lvl), c->S, sym, c->zdim, c->S, 31, c->q, st);
            c->prof.end(st);
            return rc;
        }
        nl += 2;
        c->prof.begin(CAT_ROWS_Z, st);
        int rc = bsw_logistic_2p(0, true, s, first, count, c->mu, c->zdim, c->scale, c->zdim, zend(lvl), c->S, zmeta(lvl), (int16_t *)sym, c->zdim, c->S, 31, c->q, c->scratch, c->scratch_bytes, st);
        c->prof.end(st);
        if (rc) return rc;
        serial_stream();
        c->prof.begin(CAT_POP_Z, st);
        rc = bsw_logistic_2p(1, true, s, first, count, c->mu, c->zdim, c->scale, c->zdim, zend(lvl), c->S, zmeta(lvl), (int16_t *)sym, c->zdim, c->S, 31, c->q, c->scratch, c->scratch_bytes, st);
        c->prof.end(st);
        return rc;
    }
    int push_z(int lvl, const int16_t *sym) {
        coder_stream();
        if (!c->two_phase) {
            ++nl;
            c->prof.begin(CAT_PUSH_Z, st);
            int rc = bsw_logistic_push(s, first, count, c->mu, c->zdim, c->scale, c->zdim, zend(lvl), c->S, sym, c->zdim, c->S, 31, c->q, st);
            c->prof.end(st);
            return rc;
        }
        nl += 2;
        c->prof.begin(CAT_ROWS_Z, st);
        int rc = bsw_logistic_2p(0, false, s, first, count, c->mu, c->zdim, c->scale, c->zdim, zend(lvl), c->S, zmeta(lvl), (int16_t *)sym, c->zdim, c->S, 31, c->q, c->scratch, c->scratch_bytes, st);
        c->prof.end(st);
        if (rc) return rc;
        serial_stream();
        c->prof.begin(CAT_PUSH_Z, st);
        rc = bsw_logistic_2p(1, false, s, first, count, c->mu, c->zdim, c->scale, c->zdim, zend(lvl), c->S, zmeta(lvl), (int16_t *)sym, c->zdim, c->S, 31, c->q, c->scratch, c->scratch_bytes, st);
        c->prof.end(st);
        return rc;
    }
    // p(x | z_1): ImageBins endpoints (one shared row), 8-bit quantisation (cifar_compress.py:202)
    int64_t xss() const { return c->m->d.cond_xscale ? c->xdim : 0; }
    int pop_x(int16_t *sym) {
        coder_stream();
        if (!c->two_phase) {
            ++nl;
            c->prof.begin(CAT_POP_X, st);
            int rc = bsw_logistic_pop(s, first, count, c->mu, c->xdim, c->scale, xss(), c->b->xend, 0, sym, c->xdim, 256, 31, 8, st);
            c->prof.end(st);
            return rc;
        }
        nl += 2;
        c->prof.begin(CAT_ROWS_X, st);
        int rc = bsw_logistic_2p(0, true, s, first, count, c->mu, c->xdim, c->scale, xss(), c->b->xend, 0, xmeta(), (int16_t *)sym, c->xdim, 256, 31, 8, c->scratch, c->scratch_bytes, st);
        c->prof.end(st);
        if (rc) return rc;
        serial_stream();
        c->prof.begin(CAT_POP_X, st);
        rc = bsw_logistic_2p(1, true, s, first, count, c->mu, c->xdim, c->scale, xss(), c->b->xend, 0, xmeta(), (int16_t *)sym, c->xdim, 256, 31, 8, c->scratch, c->scratch_bytes, st);
        c->prof.end(st);
        return rc;
    }
    int push_x(const int16_t *sym) {
        coder_stream();
        if (!c->two_phase) {
            ++nl;
            c->prof.begin(CAT_PUSH_X, st);
            int rc = bsw_logistic_push(s, first, count, c->mu, c->xdim, c->scale, xss(), c->b->xend, 0, sym, c->xdim, 256, 31, 8, st);
            c->prof.end(st);
            return rc;
        }
        nl += 2;
        c->prof.begin(CAT_ROWS_X, st);
        int rc = bsw_logistic_2p(0, false, s, first, count, c->mu, c->xdim, c->scale, xss(), c->b->xend, 0, xmeta(), (int16_t *)sym, c->xdim, 256, 31, 8, c->scratch, c->scratch_bytes, st);
        c->prof.end(st);
        if (rc) return rc;
        serial_stream();
        c->prof.begin(CAT_PUSH_X, st);
        rc = bsw_logistic_2p(1, false, s, first, count, c->mu, c->xdim, c->scale, xss(), c->b->xend, 0, xmeta(), (int16_t *)sym, c->xdim, 256, 31, 8, c->scratch, c->scratch_bytes, st);
        c->prof.end(st);
        return rc;
    }
    int push_prior(const int16_t *sym) {
        coder_stream();
        serial_stream();
        ++nl;
        c->prof.begin(CAT_PRIOR, st);
        int rc = bsw_ans_push_i16(s, first, count, c->priorP, c->priorC, 0, 0, sym, c->zdim, c->S, 31, st);
        c->prof.end(st);
        return rc;
    }
    int pop_prior(int16_t *sym) {
        coder_stream();
        serial_stream();
        ++nl;
        c->prof.begin(CAT_PRIOR, st);
        int rc = c->priorCoarse
            ? bsw_pop_shared_table(s, first, count, c->priorP, c->priorCoarse, c->priorFix, sym, c->zdim, c->S, 31, st)
            : bsw_ans_pop_i16(s, first, count, c->priorP, c->priorC, 0, 0, sym, c->zdim, c->S, 31, st);
        c->prof.end(st);
        return rc;
    }
};
__global__ void k_gather_x16(const int16_t *__restrict__ x, float *__restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)(((double)x[i] - 127.5) / 127.5);
}
int Ctx::gather_x16(const int16_t *x) {
    coder_stream();
    ++nl;
    int64_t n = (int64_t)count * c->xdim;
    k_gather_x16<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, c->given, n);
    BSW_LAUNCH_CHECK();
    return BSW_OK;
}
}  // namespace

static int codec_args_ok(bsw_codec *c, bsw_streams *s, int first, int count, const void *x) {
    BSW_REQUIRE(c && s && x, "bsw_codec: null argument");
    BSW_REQUIRE(first >= 0 && count >= 1 && first + count <= s->B, "bsw_codec: stream range out of bounds");
    BSW_REQUIRE(count <= c->max_batch, "bsw_codec: count exceeds the codec's max_batch");
    return BSW_OK;
}
static int ensure_zs(bsw_codec *c) {
    for (auto &p : c->zs)
        if (!p) BSW_CUDA(cudaMalloc(&p, sizeof(int16_t) * c->zdim * c->max_batch));
    return BSW_OK;
}

extern "C" int bsw_codec_encode(bsw_codec *c, bsw_streams *s, int first, int count, const uint8_t *x, int scheme, void *stream) {
    RC(codec_args_ok(c, s, first, count, x));
    Ctx k{c, s, first, count, (cudaStream_t)stream};
    k.begin((cudaStream_t)stream);
    const int nz = c->nz;
    int64_t nx = (int64_t)count * c->xdim;
    k_u8_to_i16<<<(unsigned)((nx + 255) / 256), 256, 0, k.st>>>(x, c->xsym, nx);
    BSW_LAUNCH_CHECK();
    ++k.nl;
    int16_t *zsym = c->sym[0], *ztop = c->sym[1];
    if (scheme == 0) {
        // Bit-Swap sender (cifar_compress.py:178-204)
        for (int zi = 0; zi < nz; ++zi) {
            if (zi == 0) RC(k.gather_x(x)); else RC(k.gather_z(zi - 1, zsym));           // :180
            RC(k.infer(zi));                                                                // :181
            RC(k.pop_z(zi, ztop));                                                          // :182-187
            if (zi == 0) { RC(bsw_streams_mark_rest(s, first, count, k.st)); ++k.nl; }      // :190-192 restbits (first pop of a chain only)
            RC(k.gather_z(zi, ztop));                                                       // :195
            RC(k.generate(zi));                                                             // :196
            if (zi == 0) RC(k.push_x(c->xsym)); else RC(k.push_z(zi - 1, zsym));           // :197-202
            int16_t *t = zsym; zsym = ztop; ztop = t;                                       // :204
        }
    } else {
        // BB-ANS sender (cifar_compress.py:208-242): all pops first, then all pushes
        RC(ensure_zs(c));
        for (int zi = 0; zi < nz; ++zi) {
            if (zi == 0) RC(k.gather_x(x)); else RC(k.gather_z(zi - 1, c->zs[zi - 1]));
            RC(k.infer(zi));
            RC(k.pop_z(zi, c->zs[zi]));
            if (zi == 0) { RC(bsw_streams_mark_rest(s, first, count, k.st)); ++k.nl; }      // :224-226 restbits
        }
        for (int zi = 0; zi < nz; ++zi) {
            RC(k.gather_z(zi, c->zs[zi]));
            RC(k.generate(zi));
            if (zi == 0) RC(k.push_x(c->xsym)); else RC(k.push_z(zi - 1, c->zs[zi - 1]));
        }
        zsym = c->zs[nz - 1];
    }
    RC(k.push_prior(zsym));                                                                 // :245-250
    k.end();
    c->launches = k.nl;
    return BSW_OK;
}

extern "C" int bsw_codec_decode(bsw_codec *c, bsw_streams *s, int first, int count, uint8_t *x, int scheme, void *stream) {
    RC(codec_args_ok(c, s, first, count, x));
    Ctx k{c, s, first, count, (cudaStream_t)stream};
    k.begin((cudaStream_t)stream);
    const int nz = c->nz;
    int16_t *ztop = c->sym[0], *sym = c->sym[1];
    if (scheme == 0) {
        // Bit-Swap receiver (cifar_compress.py:283-317)
        RC(k.pop_prior(ztop));                                                              // :284-289
        for (int zi = nz - 1; zi >= 0; --zi) {
            RC(k.gather_z(zi, ztop));                                                       // :296
            RC(k.generate(zi));                                                             // :297
            if (zi == 0) RC(k.pop_x(c->xsym)); else RC(k.pop_z(zi - 1, sym));               // :298-303
            if (zi == 0) RC(k.gather_x16(c->xsym)); else RC(k.gather_z(zi - 1, sym));       // :306
            RC(k.infer(zi));                                                                // :307
            RC(k.push_z(zi, ztop));                                                         // :308-313
            int16_t *t = ztop; ztop = sym; sym = t;                                         // :315
        }
    } else {
        // BB-ANS receiver (cifar_compress.py:319-352)
        RC(ensure_zs(c));
        RC(k.pop_prior(c->zs[nz - 1]));
        for (int zi = nz - 1; zi >= 0; --zi) {
            RC(k.gather_z(zi, c->zs[zi]));
            RC(k.generate(zi));
            if (zi == 0) RC(k.pop_x(c->xsym)); else RC(k.pop_z(zi - 1, c->zs[zi - 1]));
        }
        for (int zi = nz - 1; zi >= 0; --zi) {
            if (zi == 0) RC(k.gather_x16(c->xsym)); else RC(k.gather_z(zi - 1, c->zs[zi - 1]));
            RC(k.infer(zi));
            RC(k.push_z(zi, c->zs[zi]));
        }
    }
    int64_t nx = (int64_t)count * c->xdim;
    k.coder_stream();
    k_i16_to_u8<<<(unsigned)((nx + 255) / 256), 256, 0, k.st>>>(c->xsym, x, nx);
    BSW_LAUNCH_CHECK();
    k.end();
    c->launches = k.nl + 1;
    return BSW_OK;
}
