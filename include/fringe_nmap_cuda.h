/* Link-time replacement for the reference's src/nmap/nmap_cuda.h (lines 13-17): the same three
 * functions with the same C++ linkage and argument meaning, implemented by libfringe_b200.so on the
 * sm_100a kernels.  Building the reference's nmap.cpp with -DBUILD_NMAP_WITH_CUDA (the macro that guards
 * its call site, nmap.cpp:475-485) and linking libfringe_b200.so instead of its nmap_cuda.cu object is
 * the whole integration; see INTEGRATION.md, level 0.
 *   amp    float [lines*cols][bands]   amplitudes as nmap.cpp:370-381 leaves them (sorted or not)
 *   msk    uchar [lines*cols]          the reference's zeromask (non-zero = valid pixel)
 *   cnt    int   [lines*cols]          out: neighbour counts
 *   wmask  uint  [lines*cols][wtslen]  out: window bit masks, layout of include/fringe/ulongmask.hpp
 * KS2 only, like the reference's device path (nmap_cuda.cu:310). */
#ifndef FRINGE_NMAP_CUDA_H
#define FRINGE_NMAP_CUDA_H

void lockGPU();
void unlockGPU();
void nmapProcessBlock(float *amp, unsigned char *msk,
                      int cols, int lines, int bands,
                      int *cnt, unsigned int *wmask,
                      int wtslen, double pval,
                      int Nx, int Ny);

#endif
