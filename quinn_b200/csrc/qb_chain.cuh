// Chain-step building blocks shared by the CUDA-core chain kernels (qb_kernels.cu) and the tensor-core gradient
// kernels (qb_grad_tc.cu): argument structs, the Metropolis-Hastings step (mcmc.py:69-85) and the HMC / MALA step
// (hmc.py:27-70, mala.py:24-53) written against a gradient-evaluation functor.
#pragma once
#include "qb_device.cuh"

// arguments of the stand-alone evaluation kernels (kernels 1 and 2)
template <typename T> struct EvalArgs {
    const T* theta; const T* x; const T* y;
    long long xs, ys;          // per-member data strides in elements (0: x, y shared by all members)
    long long K, N; int S; long long pps;
    double* part;   // [K,S]
    T* grad;        // [K,P] (S == 1) or nullptr
    T* gpart;       // [K,S,P] (S > 1)
    double* lp;
    QbLikDev lk;
    const T* xsplit;           // hot-shape tensor-core path: x as ready-made operand tiles (qb_value_tc3.cu), or nullptr
};

template <typename T> __device__ __forceinline__ T qb_mul(T a, T b);
template <> __device__ __forceinline__ float qb_mul<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double qb_mul<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T qb_add(T a, T b);
template <> __device__ __forceinline__ float qb_add<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double qb_add<double>(double a, double b) { return __dadd_rn(a, b); }

template <typename T> struct ChainArgs {
    long long K, N;
    const T* x; const T* y;
    QbLikDev lk;
    T* theta; double* lp; long long* naccept; T* map_theta; double* map_lp;
    int rng_mode; unsigned long long seed; long long chain_offset; const T* incr; const double* unif;
    double* rec_lp; double* rec_alpha; unsigned char* rec_acc; long long rec_ld; double* rec_lp0;
    T* samples; long long store_every, n_slots;
    long long t_start, nsteps; int init_lp;
};
template <typename T> struct AmcmcArgs {
    double gamma; long long t0, tadapt; int adapt, track;
    T* xm; T* cov; T* pscale; T* chol; const T* chol_ini; int* prop_kind; T* prop;
};
template <typename T> struct HmcArgs {
    int method, L; double eps;
    T* gcur; T* mom; T* prop; T* gprop;
};

// Metropolis-Hastings accept + bookkeeping (mcmc.py:69-85); all threads hold identical scalars.
template <typename T>
__device__ __forceinline__ bool qb_mh_step(const ChainArgs<T>& c, long long k, long long s, int P, double lp_prop,
                                           double K_cur, double K_prop, T* cur, const T* prop, T* mapth,
                                           double& lp_cur, double& map_lp, long long& na) {
    const double cur_H = -lp_cur + K_cur, prop_H = -lp_prop + K_prop;
    const double mh = exp(cur_H - prop_H);                       // unclipped; inf is normal
    double u;
    if (c.rng_mode == QB_RNG_REPLAY) u = c.unif[s * c.K + k];
    else u = qb_u01(qb_rand4(c.seed, c.chain_offset + k, c.t_start + s, QB_STREAM_UNIF, 0).x);
    const bool acc = u < mh;                                      // strict <, NaN rejects
    if (acc) {
        for (int i = threadIdx.x; i < P; i += blockDim.x) cur[i] = prop[i];
        lp_cur = lp_prop;
        na += 1;
        if (lp_cur >= map_lp) {
            map_lp = lp_cur;
            for (int i = threadIdx.x; i < P; i += blockDim.x) mapth[i] = prop[i];
        }
    }
    if (threadIdx.x == 0) {
        if (c.rec_lp) c.rec_lp[k * c.rec_ld + s] = lp_cur;
        if (c.rec_alpha) c.rec_alpha[k * c.rec_ld + s] = mh;
        if (c.rec_acc) c.rec_acc[k * c.rec_ld + s] = acc ? 1 : 0;
    }
    __syncthreads();
    if (c.samples && c.store_every > 0 && (s + 1) % c.store_every == 0) {
        const long long slot = (s + 1) / c.store_every - 1;
        if (slot < c.n_slots) {
            T* dst = c.samples + (k * c.n_slots + slot) * P;
            for (int i = threadIdx.x; i < P; i += blockDim.x) dst[i] = cur[i];
        }
    }
    return acc;
}

template <typename T>
__device__ double qb_full_grad(const QbPlan& plan, const QbSmem& S, const ChainArgs<T>& c, long long k, const T* th, T* g,
                               double* lp_out) {
    const int P = plan.n_params;
    const double ssq = qb_eval_value_grad<T>(plan, S, th, c.x, c.y, 0, c.N, c.lk.inv_sigma2, g);
    double pss = 0.0;
    if (c.lk.has_prior) {
        pss = qb_prior_ss<T>(c.lk, th, k, P, S.red);
        qb_prior_grad_add<T>(c.lk, th, k, P, g);
    }
    __syncthreads();
    *lp_out = qb_lp_from(c.lk, ssq, c.N, pss, P);
    return ssq;
}

// HMC / MALA steps of chain blockIdx.x.  `eval(theta, grad)` evaluates the log-posterior (returned, block-uniform) and
// writes its gradient; it is called from ONE site (the inlined kernel-2 body is large).
template <typename T, typename GradEval>
__device__ __forceinline__ void qb_hmc_body(const ChainArgs<T>& c, const HmcArgs<T>& h, int P, double* red, GradEval& eval) {
    const long long k = blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x;
    T* cur = c.theta + k * P;
    T* mapth = c.map_theta + k * P;
    T* gcur = h.gcur + k * P;
    T* mom = h.mom + k * P;
    T* prop = h.prop + k * P;
    T* gprop = h.gprop + k * P;
    const T eps = (T)h.eps;

    // ONE gradient-evaluation call site (the inlined kernel-2 body is ~9k instructions): step s == -1 (only when
    // init_lp) evaluates the incoming state; every other step runs `nsub` sub-iterations, each a position update,
    // one evaluation and a momentum update.
    double lp_cur = 0.0, map_lp = 0.0;
    long long na = 0;
    if (!c.init_lp) { lp_cur = c.lp[k]; map_lp = c.map_lp[k]; na = c.naccept[k]; }
    const T c2 = qb_mul<T>(T(0.5), qb_mul<T>(eps, eps));
    __syncthreads();

    for (long long s = c.init_lp ? -1 : 0; s < c.nsteps; ++s) {
        const long long t = c.t_start + s;
        const bool init_step = s < 0;
        double K_cur = 0.0;
        if (!init_step) {
            // momentum draw (hmc.py:43 / mala.py:42)
            double ksum = 0.0;
            if (c.rng_mode == QB_RNG_REPLAY) {
                const T* p0 = c.incr + (s * c.K + k) * P;
                for (int i = tid; i < P; i += nt) { const T v = p0[i]; mom[i] = v; ksum += (double)v * (double)v; }
            } else {
                const long long chain = c.chain_offset + k;
                for (int i4 = tid; i4 * 4 < P; i4 += nt) {
                    T z[4];
                    qb_normal4(qb_rand4(c.seed, chain, t, QB_STREAM_INCR, (uint32_t)i4), z);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = i4 * 4 + q;
                        if (i < P) { mom[i] = z[q]; ksum += (double)z[q] * (double)z[q]; }
                    }
                }
            }
            K_cur = qb_block_sum(ksum, red) / 2.0;
            if (h.method == 0) {
                // first half step of the leapfrog (hmc.py:48); each thread only touches its own elements
                for (int i = tid; i < P; i += nt) {
                    mom[i] = qb_add<T>(mom[i], qb_mul<T>(eps, gcur[i]) / T(2));
                    prop[i] = cur[i];
                }
            }
        }
        const int nsub = init_step ? 1 : (h.method == 0 ? h.L : 1);
        double lp_eval = 0.0;
        for (int jj = 0; jj < nsub; ++jj) {
            if (!init_step) {
                if (h.method == 0) {          // hmc.py:52
                    for (int i = tid; i < P; i += nt) prop[i] = qb_add<T>(prop[i], qb_mul<T>(eps, mom[i]));
                } else {                      // mala.py:45
                    for (int i = tid; i < P; i += nt)
                        prop[i] = qb_add<T>(cur[i], qb_add<T>(qb_mul<T>(c2, gcur[i]), qb_mul<T>(eps, mom[i])));
                }
            }
            __syncthreads();
            lp_eval = eval(init_step ? cur : prop, init_step ? gcur : gprop);
            if (!init_step) {
                if (h.method == 0) {          // hmc.py:57 / :60
                    if (jj != nsub - 1) {
                        for (int i = tid; i < P; i += nt) mom[i] = qb_add<T>(mom[i], qb_mul<T>(eps, gprop[i]));
                    } else {
                        for (int i = tid; i < P; i += nt) mom[i] = qb_add<T>(mom[i], qb_mul<T>(eps, gprop[i]) / T(2));
                    }
                } else {                      // mala.py:50
                    for (int i = tid; i < P; i += nt)
                        mom[i] = qb_add<T>(mom[i], qb_mul<T>(eps, qb_add<T>(gcur[i], gprop[i])) / T(2));
                }
            }
        }
        if (init_step) {
            lp_cur = lp_eval; map_lp = lp_eval; na = 0;
            if (tid == 0 && c.rec_lp0) c.rec_lp0[k] = lp_cur;
            for (int i = tid; i < P; i += nt) mapth[i] = cur[i];
        } else {
            double k2 = 0.0;
            for (int i = tid; i < P; i += nt) { const double v = (double)mom[i]; k2 += v * v; }
            const double K_prop = qb_block_sum(k2, red) / 2.0;
            const bool acc = qb_mh_step<T>(c, k, s, P, lp_eval, K_cur, K_prop, cur, prop, mapth, lp_cur, map_lp, na);
            if (acc) for (int i = tid; i < P; i += nt) gcur[i] = gprop[i];
        }
        __syncthreads();
    }
    if (tid == 0) { c.lp[k] = lp_cur; c.map_lp[k] = map_lp; c.naccept[k] = na; }
}

