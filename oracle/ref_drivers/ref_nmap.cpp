// TEST INFRASTRUCTURE ONLY -- src/nmap/nmap.cpp compiled as it stands (see common.hpp)
#include "common.hpp"
#define main ref_nmap_main
#include "nmap.cpp"
#undef main
extern "C" int ref_nmap(const char* input, const char* wts, const char* count, const char* mask, int Nx, int Ny,
                        const char* method, double prob, int memsize, int blocksize) {
    nmapOptions o;
    o.inputDS = input; o.wtsDS = wts; o.ncountDS = count; o.maskDS = mask ? mask : "";
    o.Nx = Nx; o.Ny = Ny; o.method = method; o.prob = prob; o.memsize = memsize; o.blocksize = blocksize; o.noGPU = true;
    return nmap_process(&o);
}
