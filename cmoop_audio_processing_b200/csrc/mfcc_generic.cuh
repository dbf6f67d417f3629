// Generic (any power-of-two n_fft, optional centring) log-mel / MFCC path; see mfcc_generic.cu.
#pragma once
#include "common.cuh"

namespace cmoop {
struct GenericMfcc;
int generic_mfcc_create(const cmoop_mfcc_config* cfg, GenericMfcc** out);
void generic_mfcc_destroy(GenericMfcc* g);
int generic_mfcc_n_out(const GenericMfcc* g);
int generic_mfcc_n_frames(const GenericMfcc* g, int n_samples);
int generic_mfcc_set_standardise(GenericMfcc* g, const float* mean, const float* scale);
int generic_mfcc_fwd(GenericMfcc* g, const float* wave, int64_t n_clips, int n_samples, float* out, void* stream);
}  // namespace cmoop
