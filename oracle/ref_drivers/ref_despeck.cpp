// TEST INFRASTRUCTURE ONLY -- src/despeck/despeck.cpp compiled as it stands (see common.hpp)
#include "common.hpp"
#define main ref_despeck_main
#include "despeck.cpp"
#undef main
extern "C" int ref_despeck(const char* input, const char* wts, const char* output, int Nx, int Ny, int band1, int band2,
                           int coherence, int memsize, int blocksize) {
    despeckOptions o;
    o.inputDS = input; o.wtsDS = wts; o.outputDS = output; o.Nx = Nx; o.Ny = Ny; o.ibands[0] = band1; o.ibands[1] = band2;
    o.computeCoherence = coherence != 0; o.memsize = memsize; o.blocksize = blocksize;
    return despeck_process(&o);
}
