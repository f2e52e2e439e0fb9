// Launch plan shared by host and device: how one thread block tiles a small MLP.
//
// Data layout inside a block (see DESIGN.md "Data layout"):
//   * activations live in shared memory as rows = units, columns = points of the current tile of TM
//     data points: A[unit * lda + point], lda = TM + LDPAD so that consecutive rows start 4 banks apart
//     (conflict-free 128-bit row accesses for 8 consecutive units and for 8 consecutive point groups);
//   * every layer's weights are staged once per chain: Wt_l[i][col] (i = input unit, col = permuted
//     output unit) for the forward GEMM, Wr_l[j][col] (j = output unit, col = permuted input unit) for
//     the back-propagation GEMM; "permuted" means column ug*TU+u holds true unit ug + UG*u so that the
//     TU units of one thread tile are contiguous in the weight row (one 128-bit load) but interleaved in
//     the activation rows (conflict-free stores).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include "quinn_b200.h"

enum { QB_MODE_GEMM = 0, QB_MODE_DOT = 1 };

struct QbLayerPlan {
    int n_in, n_out, n_in_pad, n_out_pad;
    int w_off, b_off;             // offsets in the flat parameter vector (b_off < 0: none)
    int wt_off, wr_off, bias_off; // element offsets in the shared weight area (wr_off < 0: absent)
    int row_in, row_out;          // row offsets of this layer's input / output activations (grad kernel)
    int act, mode, nj, dw_chunks;
    int has_res, ug_shift;        // ug_shift: log2(n_out_pad/TU) if that is a power of two, else -1
    int ugi_shift, ig_shift, c_shift, pad2_;   // same for n_in_pad/TU, ceil(n_in/4), dw_chunks
    double res_step;
    int n_terms, w_stride, b_stride, pad3_;   // polynomial-in-depth weights (qb_layer_t): W = sum_m coef[m]*theta[w_off + m*w_stride ..]
    double coef[QB_MAX_TERMS];
};

struct QbPlan {
    int n_layers, in_dim, out_dim, n_params, final_exp;
    int TM, lda, nthreads;
    int inplace;        // value kernel: layers overwrite their input buffer
    int buf_rows;       // value kernel: rows per activation buffer
    int total_rows;     // grad kernel: rows of all kept activations (+ residual pass-through buffer)
    int d_row;          // grad kernel: first row of the residual pass-through buffer, or -1
    int w_elems;        // shared weight area, elements
    int want_grad;
    int has_res;
    int elem_size;
    int ws, WP;         // value kernel: warp-synchronous mode, points owned by one warp
    int tpg;            // gradient kernel: points per thread tile (4 or 8)
    int fuse_tail;      // value kernel: last (narrow linear) layer folded into the epilogue of the last hidden layer
    long long smem_bytes;
    QbLayerPlan L[QB_MAX_LAYERS];
};
