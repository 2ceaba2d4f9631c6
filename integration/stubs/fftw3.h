/* Build-only stand-in for <fftw3.h>: the reference's generic_functions_factories.cpp includes its FFTW DFT header
 * unconditionally although the FFTW processor itself is compiled only with ENABLE_FFTW (not defined here: the generic DFT
 * is used). Only the names that header mentions are declared; nothing here is ever called. */
#ifndef PDC_INTEGRATION_FFTW3_STUB_H
#define PDC_INTEGRATION_FFTW3_STUB_H
typedef struct fftwf_plan_s* fftwf_plan;
typedef float                fftwf_complex[2];
#define FFTW_MEASURE (0U)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_ESTIMATE (1U << 6)
#endif
