// TEST INFRASTRUCTURE ONLY -- src/calamp/calamp.cpp compiled as it stands (see common.hpp)
#include "common.hpp"
#define main ref_calamp_main
#include "calamp.cpp"
#undef main
// constants[b] = the amplitudeConstant calamp.cpp:264 attaches to band b+1 of its output VRT (the shim keeps the copy in memory)
extern "C" int ref_calamp(const char* input, const char* mask, const char* output, double default_value, int apply_sqrt,
                          int memsize, int blocksize, int nbands, double* constants) {
    calampOptions o;
    o.inputDS = input; o.maskDS = mask ? mask : ""; o.outputDS = output; o.defaultValue = default_value;
    o.applySqrt = apply_sqrt != 0; o.memsize = memsize; o.blocksize = blocksize;
    GDALDriver::gdal_shim_last_copy() = nullptr;
    const int rc = calamp_process(&o);
    GDALDataset* copy = GDALDriver::gdal_shim_last_copy();
    for (int b = 0; b < nbands; ++b) {
        const char* v = (copy && b < copy->nb) ? copy->bands[b].GetMetadataItem("amplitudeConstant", "slc") : nullptr;
        constants[b] = v ? std::atof(v) : -1.0;
    }
    delete copy;
    GDALDriver::gdal_shim_last_copy() = nullptr;
    return rc;
}
