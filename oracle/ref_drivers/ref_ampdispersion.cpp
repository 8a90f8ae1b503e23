// TEST INFRASTRUCTURE ONLY -- src/ampdispersion/ampdispersion.cpp compiled as it stands (see common.hpp)
#include "common.hpp"
#define main ref_ampdispersion_main
#include "ampdispersion.cpp"
#undef main
extern "C" int ref_ampdispersion(const char* input, const char* da, const char* meanamp, int refband, int memsize, int blocksize) {
    ampdispersionOptions o;
    o.inputDS = input; o.daDS = da; o.meanampDS = meanamp; o.refband = refband; o.memsize = memsize; o.blocksize = blocksize;
    return ampdispersion_process(&o);
}
