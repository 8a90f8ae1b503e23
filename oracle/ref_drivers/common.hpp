// TEST INFRASTRUCTURE ONLY -- shared preamble of the reference-driver builds (oracle/ref_drivers/*.cpp).
//
// Each of those files includes ONE of the reference's block drivers textually and unmodified from /root/reference
// (the way the reference's own Cython modules do, nmaplib.pyx:24-25) and exports a C entry that fills the driver's
// option struct and calls its xxx_process().  GDAL and Armadillo come from oracle/shims/ (storage and file I/O only).
//
// One reference helper header is replaced here: include/fringe/fringe_common.hpp.  Its numberOfThreads() aborts unless at
// least two OpenMP threads run, and with two or more threads nmap.cpp's pair loop updates count(qq) / wts(qq) without
// synchronisation (nmap.cpp:464-468) -- the result would not be reproducible.  These builds therefore compile WITHOUT
// OpenMP (the pragmas are ignored: one thread, the race-free reading of the loop) and provide the three helpers the
// drivers take from that header.
//
// -DFRINGE_REF_OPENMP (the *_omp.so builds, used only to TIME the reference's drivers on all host cores in bench.py's
// reference arm): nothing is replaced, the drivers are compiled with -fopenmp exactly as the reference builds them.
#pragma once
#ifndef FRINGE_REF_OPENMP
#define FRINGE_COMMON_H
#include <sys/time.h>

#include <iostream>
#include <string>

inline double getWallTime() {
    struct timeval t;
    if (gettimeofday(&t, NULL)) return 0;
    return (double)t.tv_sec + (double)t.tv_usec * .000001;
}
inline int numberOfThreads() { return 1; }
inline int omp_get_thread_num() { return 0; }
#endif  // FRINGE_REF_OPENMP
