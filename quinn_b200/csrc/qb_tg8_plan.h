// Launch plan of the fp16-split tensor-core gradient kernel (qb_tg8.cuh), shared by host and device.
#pragma once

struct QbTg8Plan {
    int in_dim, n_params, h, hr;                          // h: width of the kernel instance, hr: the net's own width (32 runs padded on 64)
    int w0_off, b0_off, w1_off, b1_off, wl_off, bl_off;   // offsets in theta (b*_off < 0: no bias)
    int w_img, w0_img, a_img, z_img, x_img;               // byte offsets: hi image, then lo image (X: two tiles of hi | lo)
    int fl_base, b1, wl, bl, sc;                          // float area (byte offset) and float indices in it
    int ybuf;                                             // byte offset: [H/32][128] partial outputs
    int tmem_cols, nthreads, smem_bytes;
};
