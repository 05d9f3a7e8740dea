// Shared between the two inter-prediction kernels (inter.cu: generic CTA-per-record kernel for any
// bit depth / alignment; inter_warp.cu: the 10-bit warp-per-record kernel the batched entry uses).
#pragma once
#include "common.cuh"

struct InterK {
    const pel *ref[3];
    pel       *dst[3];
    int        rp[3], dp[3];
    long long  rb[3], db[3];
    int        w, h, bd, planes;
    const VVCCudaPB   *pbs;
    int                n;
    const VVCCudaWP   *wp;
    const VVCCudaProf *prof;
    VVCCudaDmvrOut    *dmvr_out;
};

// inter_warp.cu: requires bd == 10, 4:2:0 or 4:0:0, and 16-byte aligned planes / pitches
int vvc_inter_launch_warp(VVCCudaCtx *ctx, const InterK &p);
