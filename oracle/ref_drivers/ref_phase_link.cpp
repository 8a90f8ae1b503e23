// TEST INFRASTRUCTURE ONLY -- src/phase_link/phase_link.cpp compiled as it stands (see common.hpp)
#include "common.hpp"
#define main ref_phase_link_main
#include "phase_link.cpp"
#undef main
extern "C" int ref_phase_link(const char* input, const char* wts, const char* out_folder, const char* comp_folder,
                              const char* comp_name, int Nx, int Ny, const char* method, int bandwidth, int mini_stack_count,
                              int min_neighbors, int memsize, int blocksize) {
    evdOptions o;
    o.inputDS = input; o.wtsDS = wts; o.outputFolder = out_folder; o.outputCompressedSlcFolder = comp_folder; o.compSlc = comp_name;
    o.Nx = Nx; o.Ny = Ny; o.method = method; o.bandWidth = bandwidth; o.miniStackCount = mini_stack_count;
    o.minNeighbors = min_neighbors; o.memsize = memsize; o.blocksize = blocksize;
    return evd_process(&o);
}
