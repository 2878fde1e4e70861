// nms_train.cuh -- parameter block of the training-step kernel (nms_train.cu), shared with the launcher.
#pragma once
#include "nms_common.cuh"

#include <cstddef>
#include <cstdint>

struct TrainParams {
    int M, N, E, z, NZ, EZ, MZ;
    const int *row, *col, *shift, *row_ptr, *col_ptr, *col_edge;   // device, E(C) order
    int qms;
    float qmagic, qmax, clip;
    int sharing0, sharing1, sharing2, wc, wu, wv;
    const float *w;            // device [T*wc | T*wu | T*wv]
    int off_cn, off_ucn, off_vn;
    int T, t_lo, loss_type, target_nz, B;
    const float *coef;         // device [T]: eta^(T-1-t) / sum, 0 below t_lo
    const float *llr;          // [B, NZ]
    float *hist;               // [B, T+1, EZ]
    double *loss;              // [1], accumulated
    float *grad;               // [T*wc | T*wu | T*wv], accumulated
    float *app_out;            // optional [T, B, NZ]
};

size_t nms_train_smem_bytes(const TrainParams &P);
cudaError_t nms_launch_train(const TrainParams &P, cudaStream_t st);
