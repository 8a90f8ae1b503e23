// Block drivers: nmap_process / evd_process / phase_link_process.
//
// Same contract as the reference's drivers (src/nmap/nmap.cpp:18-613, src/evd/evd.cpp:16-888,
// src/phase_link/phase_link.cpp:17-755): read the stack VRT in overlapping blocks of lines
// (block height from `memsize`, Ny-line overlap, interior lines written), produce the same
// ENVI rasters with the same names, types, interleave and metadata, return the same error
// codes.  What differs is where the pixels are computed: every block goes through the C ABI
// of libfringe_b200.so (fringe_nmap_block / fringe_evd_block) from pinned host buffers, and
// blocks are dealt to all visible GPUs (two worker threads per GPU, each with its own context + buffer set; blocks
// are independent, so no inter-GPU traffic).  There is no CPU compute path: if the device
// library reports an error the driver returns 200 + that status.
#include <strings.h>
#include <sys/stat.h>

#include <atomic>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <mutex>
#include <thread>

#include "../../../include/fringe_b200.h"
#include "options.hpp"
#include "raster_io.hpp"

using fringe_host::EnviWriter;
using fringe_host::Raster;

namespace {

struct Block { int yoff, inysize, first, nwrite; };

// The reference's streaming schedule (nmap.cpp:302-573 / evd.cpp:399-872).
std::vector<Block> make_schedule(int rows, int blockysize, int Ny) {
    std::vector<Block> out;
    int yoff = 0, blockcount = 0;
    while (yoff < rows) {
        ++blockcount;
        int inysize = blockysize;
        if (yoff + inysize > rows) inysize = rows - yoff;
        const bool last = (yoff + blockysize) >= rows;
        int first, nwrite, rollback;
        if (blockcount == 1) {
            first = 0; nwrite = inysize - Ny; rollback = Ny;
            if (last) { nwrite = inysize; rollback = 0; }
        } else if (last) { first = Ny; nwrite = inysize - Ny; rollback = 0; }
        else { first = Ny; nwrite = inysize - 2 * Ny; rollback = 0; }
        out.push_back({yoff, inysize, first, nwrite});
        if (!last) yoff += nwrite - rollback; else yoff = rows;
    }
    return out;
}

int block_height(int memsize, int cols, int blocksize, int denom, int rows, int Ny) {
    int boxes = int((memsize * 1.0e6) / cols) / (blocksize * denom);
    int h = boxes * blocksize;
    if (h < blocksize) h = blocksize;
    if (h > rows) h = rows;
    // a block must be taller than its two halos or the schedule cannot advance
    if (h < rows && h <= 2 * Ny) h = std::min(rows, 2 * Ny + blocksize);
    return h;
}

// Block workers: two per GPU (each with its own context, stream and pinned block), so that one worker's file reads
// and writes overlap the other's time on the device; the memory budget is shared between all workers.
int workers_for(int ngpu) { return 2 * ngpu; }

int visible_gpus() {
    int n = 0;
    if (fringe_device_count(&n) != FRINGE_OK) return 0;
    if (const char* e = std::getenv("FRINGE_NUM_GPUS")) { const int m = std::atoi(e); if (m > 0 && m < n) n = m; }
    return n;
}

struct Pinned {
    void* p = nullptr;
    bool alloc(size_t bytes) { return fringe_host_alloc(&p, bytes ? bytes : 1) == FRINGE_OK; }
    ~Pinned() { if (p) fringe_host_free(p); }
};

}  // namespace

// =====================================================================================================
int nmap_process(nmapOptions* opts) {
    opts->print();
    int method;
    if (opts->method.compare("KS2") == 0) { std::cout << "Using Kolmogorov-Smirnov 2-sample test\n"; method = FRINGE_NMAP_KS2; }
    else if (opts->method.compare("AD2") == 0) { std::cout << "Using Anderson-Darling 2-sample test \n"; method = FRINGE_NMAP_AD2; }
    else {
        std::cout << "Statistics method can be KS2 or AD2\nUnknown method: " << opts->method << "\nReturning with non-zero error code \n";
        return 1;
    }
    const int Nx = opts->Nx, Ny = opts->Ny;
    const int nulong = fringe_nulong(Nx, Ny);
    std::cout << "Number of uint32 bytes for mask: " << nulong << "\n";

    Raster in;
    if (!in.open(opts->inputDS)) {
        std::cout << "Cannot open stack file { " << opts->inputDS << " } for reading: " << in.error << "\nExiting with error code .... (102) \n";
        return 102;
    }
    if (!in.is_cfloat32_stack()) {
        std::cout << "Input stack { " << opts->inputDS << " } is not a VRT with one flat CFloat32 file per band\nExiting with error code .... (102) \n";
        return 102;
    }
    const int cols = in.cols, rows = in.rows, nbands = in.count();
    std::cout << "Number of rows  = " << rows << "\nNumber of cols  = " << cols << "\nNumber of bands = " << nbands << "\n";
    Raster msk;
    bool have_mask = false;
    if (!opts->maskDS.empty() && strcasecmp(opts->maskDS.c_str(), "None") != 0) {
        if (!msk.open(opts->maskDS)) {
            std::cout << "Cannot open mask file { " << opts->maskDS << " } for reading. \nExiting with error code .... (102) \n";
            return 102;
        }
        int code = 0;
        if (msk.cols != cols) { std::cout << "Mask file width does not match stack size width \n"; code = 104; }
        if (msk.rows != rows) { std::cout << "Mask file length does not match stack size length \n"; code = 105; }
        if (msk.count() != 1) { std::cout << "Mask file has more than one band \n"; code = 106; }
        if (code) { std::cout << "Exiting with error code .... (" << code << ")\n"; return code; }
        have_mask = true;
    }

    const int ngpu = visible_gpus();
    if (ngpu <= 0) { std::cout << "No CUDA device available and there is no CPU path.\n"; return 200 + FRINGE_ERR_NO_DEVICE; }
    std::cout << "Executing on " << ngpu << " GPU(s)\n";

    // every worker (one per GPU) holds its own pinned block: the memory budget is shared between them
    const int blockysize = block_height(std::max(1, opts->memsize / workers_for(ngpu)), cols, opts->blocksize, 4 * (nbands + 2 + nulong), rows, Ny);
    std::cout << "Block size = " << blockysize << " lines \n";
    const std::vector<Block> sched = make_schedule(rows, blockysize, Ny);
    std::cout << "Total number of blocks to process: " << sched.size() << "\n";

    // amplitude calibration constants (nmap.cpp:204-233)
    std::vector<double> alpha(nbands, 1.0);
    for (int b = 0; b < nbands; ++b) {
        double c = 0.0;
        auto it = in.bands.empty() ? std::map<std::string, std::string>::const_iterator() : in.bands[b].md_slc.find("amplitudeConstant");
        if (!in.bands.empty() && it != in.bands[b].md_slc.end()) c = std::atof(it->second.c_str());
        if (c <= 0.0) { c = 1.0; std::cout << "No calibration constant found for band " << b + 1 << ". Setting to 1.0. \n"; }
        alpha[b] = c;
    }
    {
        const double norm = alpha[0];
        for (auto& a : alpha) a /= norm;
        alpha[0] = 1.0;
    }
    bool unit_alpha = true;
    for (double a : alpha) unit_alpha = unit_alpha && (a == 1.0);

    EnviWriter wcount, wwts;
    if (!wcount.create(opts->ncountDS, cols, rows, 1, 2)) {
        std::cout << "Could not create count dataset {" << opts->ncountDS << "} \nExiting with non-zero error code ... 104 \n";
        return 104;
    }
    if (!wwts.create(opts->wtsDS, cols, rows, nulong, 13)) {
        std::cout << "Could not create weights dataset {" << opts->wtsDS << "}\nExiting with non-zero error code ... 105 \n";
        return 105;
    }

    std::atomic<size_t> next(0);
    std::atomic<int> rc(0);
    std::mutex log_mu;
    auto worker = [&](int dev) {
        fringe_ctx* ctx = nullptr;
        if (fringe_create(dev, &ctx) != FRINGE_OK) { rc = 200 + FRINGE_ERR_NO_DEVICE; return; }
        const size_t bp = (size_t)cols * blockysize;
        Pinned slc, mask, count, wts;
        std::vector<int16_t> c16(bp);
        std::vector<char> mask_scratch;
        if (!slc.alloc(bp * nbands * 8) || !mask.alloc(bp) || !count.alloc(bp * 4) || !wts.alloc(bp * nulong * 4)) {
            rc = 200 + FRINGE_ERR_MEMORY; fringe_destroy(ctx); return;
        }
        for (size_t i = next++; i < sched.size() && rc == 0; i = next++) {
            const Block& b = sched[i];
            const size_t np = (size_t)cols * b.inysize;
            bool ok = true;
            for (int band = 0; band < nbands && ok; ++band)
                ok = in.read_band_lines(band, b.yoff, b.inysize, (char*)slc.p + (size_t)band * np * 8, 8);
            if (ok && have_mask) ok = msk.read_mask_lines(b.yoff, b.inysize, (uint8_t*)mask.p, mask_scratch);
            if (!ok) {
                std::lock_guard<std::mutex> g(log_mu);
                std::cout << "Error reading data at line " << b.yoff << "\nExiting with error code .... (108) \n";
                rc = 108; break;
            }
            const int st = fringe_nmap_block(ctx, (const float*)slc.p, have_mask ? (const uint8_t*)mask.p : nullptr,
                                             unit_alpha ? nullptr : alpha.data(), cols, b.inysize, nbands, Nx, Ny,
                                             method, opts->prob, (int32_t*)count.p, (uint32_t*)wts.p);
            if (st != FRINGE_OK) {
                std::lock_guard<std::mutex> g(log_mu);
                std::cout << "Device error: " << fringe_last_error(ctx) << "\n";
                rc = 200 + st; break;
            }
            const int32_t* c32 = (const int32_t*)count.p + (size_t)b.first * cols;
            const size_t nw = (size_t)b.nwrite * cols;
            for (size_t k = 0; k < nw; ++k) c16[k] = (int16_t)std::min(c32[k], 32767);
            const bool w1 = wcount.write_lines(b.yoff + b.first, b.nwrite, c16.data());
            const bool w2 = wwts.write_lines(b.yoff + b.first, b.nwrite, (const uint32_t*)wts.p + (size_t)b.first * cols * nulong);
            if (!w1 || !w2) {
                std::lock_guard<std::mutex> g(log_mu);
                std::cout << "Error writing wts data at line " << b.yoff << "\nExiting with error code .... (111) \n";
                rc = 111; break;
            }
        }
        fringe_destroy(ctx);
    };
    {
        std::vector<std::thread> th;
        const int nw = (int)std::min<size_t>(workers_for(ngpu), sched.size());
        for (int d = 0; d < nw; ++d) th.emplace_back(worker, d % ngpu);
        for (auto& t : th) t.join();
    }
    if (rc != 0) return rc;
    for (EnviWriter* w : {&wcount, &wwts}) {       // nmap.cpp:588-591
        w->set_metadata("HALFWINDOWX", std::to_string(Nx));
        w->set_metadata("HALFWINDOWY", std::to_string(Ny));
        w->close_file();
    }
    return 0;
}

// =====================================================================================================
static int evd_driver(evdOptions* opts, int variant) {
    opts->print();
    int method;
    if (opts->method.compare("MLE") == 0) method = FRINGE_EVD_MLE;
    else if (opts->method.compare("STBAS") == 0) method = FRINGE_EVD_STBAS;
    else method = FRINGE_EVD_EVD;      // the reference treats anything else as EVD (evd.cpp:507-510,689)
    const int Nx = opts->Nx, Ny = opts->Ny;
    const int nulong = fringe_nulong(Nx, Ny);

    Raster in;
    if (!in.open(opts->inputDS)) {
        std::cout << "Cannot open stack file { " << opts->inputDS << " } for reading: " << in.error << "\n";
        return 102;
    }
    if (!in.is_cfloat32_stack()) {
        std::cout << "Input stack { " << opts->inputDS << " } is not a VRT with one flat CFloat32 file per band\nExiting with error code .... (102) \n";
        return 102;
    }
    const int cols = in.cols, rows = in.rows, nbands = in.count();
    std::cout << "Number of rows  = " << rows << "\nNumber of cols  = " << cols << "\nNumber of bands = " << nbands << "\n";
    if (method == FRINGE_EVD_STBAS) {
        if (opts->bandWidth <= 0) { std::cout << "Requested STBAS but no bandwidth provided \n"; return 101; }
        if (opts->bandWidth >= nbands - 1) {
            std::cout << "Requested STBAS bandwidth " << opts->bandWidth << " is larger than full bandwidth " << (nbands - 1) << "\n";
            return 101;
        }
    }
    Raster wts;
    if (!wts.open(opts->wtsDS)) { std::cout << "Could not open weights dataset {" << opts->wtsDS << "}\nExiting with non-zero error code ... 105 \n"; return 105; }
    {
        int code = 0;
        if (wts.cols != cols) { std::cout << "Width mismatch between input dataset and weight dataset\n"; code = 106; }
        if (wts.rows != rows) { std::cout << "Length mismatch between input dataset and weight dataset \n"; code = 107; }
        if (wts.count() != nulong) { std::cout << "Number of bands mismatch for weights and window size \n"; code = 108; }
        auto geti = [&](const char* k) { auto it = wts.envi.fields.find(k); return it == wts.envi.fields.end() ? 0 : std::atoi(it->second.c_str()); };
        const int inNx = geti("halfwindowx"), inNy = geti("halfwindowy");
        if (inNx != Nx) { std::cout << "Half window size x of wts is different from input. \n"; code = 109; }
        if (inNx == 0) { std::cout << "No non-zero metadata item called HALFWINDOWX \n"; code = 109; }
        if (inNy != Ny) { std::cout << "Half window size y of wts is different from input. \n"; code = 110; }
        if (inNy == 0) { std::cout << "No non-zero metadata item called HALFWINDOWY \n"; code = 110; }
        if (code) { std::cout << "Exiting with error code ....(" << code << ")\n"; return code; }
        if (!wts.interleaved || wts.envi.data_type != 13) { std::cout << "Weights dataset must be a UInt32 ENVI BIP raster\n"; return 105; }
    }
    if (opts->miniStackCount < 1 || opts->miniStackCount > nbands) { std::cout << "miniStackCount outside [1, bands]\n"; return 200 + FRINGE_ERR_ARGUMENT; }
    if (nbands > fringe_evd_max_bands(method, variant)) { std::cout << "Too many bands for the device kernels\n"; return 200 + FRINGE_ERR_UNSUPPORTED; }
    const int ngpu = visible_gpus();
    if (ngpu <= 0) { std::cout << "No CUDA device available and there is no CPU path.\n"; return 200 + FRINGE_ERR_NO_DEVICE; }

    const int blockysize = block_height(std::max(1, opts->memsize / workers_for(ngpu)), cols, opts->blocksize, nbands * 20 + 4 + nulong, rows, Ny);
    std::cout << "Block size = " << blockysize << " lines \n";
    const std::vector<Block> sched = make_schedule(rows, blockysize, Ny);
    std::cout << "Total number of blocks to process: " << sched.size() << "\n";

    std::vector<std::string> dates(nbands);
    for (int b = 0; b < nbands; ++b) {
        auto it = in.bands[b].md_slc.find("Date");
        if (it == in.bands[b].md_slc.end() || it->second.size() != 8) {
            std::cout << "Band " << b + 1 << " does not appear to have Date information in slc metadata domain \n";
            return 112;
        }
        dates[b] = it->second;
    }
    struct stat st;
    if (::stat(opts->outputFolder.c_str(), &st) == 0) {
        if (S_ISDIR(st.st_mode)) {
            std::cout << "Output folder : " << opts->outputFolder << " already exists \nReturning without processing. Clean up output folder and rerun.. \n";
            return 113;
        }
        std::cout << opts->outputFolder << " already exists and appears to be a file on disk. \n";
        return 114;
    }
    if (::mkdir(opts->outputFolder.c_str(), 0777) != 0) { std::cout << "Could not create output folder: " << opts->outputFolder << "\n"; return 115; }
    std::cout << "Created output folder: " << opts->outputFolder << "\n";

    std::vector<EnviWriter> wout(nbands);
    for (int b = 0; b < nbands; ++b) {
        const std::string fname = opts->outputFolder + "/" + dates[b] + ".slc";
        if (!wout[b].create(fname, cols, rows, 1, 6)) { std::cout << "Could not create output SLC: " << fname << "\nExiting with non-zero error code ... 116 \n"; return 116; }
    }
    EnviWriter wcorr, wcomp;
    if (!wcorr.create(opts->outputFolder + "/tcorr.bin", cols, rows, 1, 4)) { std::cout << "Could not create temporal correlation file\nExiting with non-zero error code ... 117 \n"; return 117; }
    {
        const std::string folder = opts->outputCompressedSlcFolder.empty() ? opts->outputFolder : opts->outputCompressedSlcFolder;
        const std::string name = opts->compSlc.empty() ? std::string("compslc.bin") : opts->compSlc;
        if (!wcomp.create(folder + "/" + name, cols, rows, 1, 6)) { std::cout << "Could not create compressed SLC file: " << folder + "/" + name << "\nExiting with non-zero error code ... 117 \n"; return 117; }
    }

    std::atomic<size_t> next(0);
    std::atomic<int> rc(0);
    std::mutex log_mu;
    auto worker = [&](int dev) {
        fringe_ctx* ctx = nullptr;
        if (fringe_create(dev, &ctx) != FRINGE_OK) { rc = 200 + FRINGE_ERR_NO_DEVICE; return; }
        const size_t bp = (size_t)cols * blockysize;
        Pinned slc, wbuf, out, tcorr, comp;
        if (!slc.alloc(bp * nbands * 8) || !wbuf.alloc(bp * nulong * 4) || !out.alloc(bp * nbands * 8) ||
            !tcorr.alloc(bp * 4) || !comp.alloc(bp * 8)) { rc = 200 + FRINGE_ERR_MEMORY; fringe_destroy(ctx); return; }
        for (size_t i = next++; i < sched.size() && rc == 0; i = next++) {
            const Block& b = sched[i];
            const size_t np = (size_t)cols * b.inysize;
            bool ok = true;
            for (int band = 0; band < nbands && ok; ++band)
                ok = in.read_band_lines(band, b.yoff, b.inysize, (char*)slc.p + (size_t)band * np * 8, 8);
            if (!ok) { std::lock_guard<std::mutex> g(log_mu); std::cout << "Error reading data at line " << b.yoff << "\n"; rc = 118; break; }
            if (!wts.read_interleaved_lines(b.yoff, b.inysize, wbuf.p)) {
                std::lock_guard<std::mutex> g(log_mu); std::cout << "Error reading weights at line " << b.yoff << "\n"; rc = 119; break;
            }
            const int stt = fringe_evd_block(ctx, (const float*)slc.p, (const uint32_t*)wbuf.p, cols, b.inysize, nbands, Nx, Ny,
                                             b.first, b.nwrite, method, opts->bandWidth, opts->miniStackCount, variant,
                                             opts->minNeighbors, (float*)out.p, (float*)tcorr.p, (float*)comp.p);
            if (stt != FRINGE_OK) {
                std::lock_guard<std::mutex> g(log_mu);
                std::cout << "Device error: " << fringe_last_error(ctx) << "\n";
                rc = 200 + stt; break;
            }
            const size_t off = (size_t)b.first * cols;
            for (int band = 0; band < nbands && ok; ++band)
                ok = wout[band].write_lines(b.yoff + b.first, b.nwrite, (const char*)out.p + ((size_t)band * np + off) * 8);
            if (!ok) { rc = 120; break; }
            if (!wcorr.write_lines(b.yoff + b.first, b.nwrite, (const float*)tcorr.p + off) ||
                !wcomp.write_lines(b.yoff + b.first, b.nwrite, (const char*)comp.p + off * 8)) { rc = 121; break; }
        }
        fringe_destroy(ctx);
    };
    {
        std::vector<std::thread> th;
        const int nw = (int)std::min<size_t>(workers_for(ngpu), sched.size());
        for (int d = 0; d < nw; ++d) th.emplace_back(worker, d % ngpu);
        for (auto& t : th) t.join();
    }
    if (rc != 0) return rc;
    for (auto& w : wout) w.close_file();
    wcorr.close_file();
    wcomp.close_file();
    return 0;
}

int evd_process(evdOptions* opts) { return evd_driver(opts, FRINGE_VARIANT_EVD); }
int phase_link_process(evdOptions* opts) { return evd_driver(opts, FRINGE_VARIANT_PHASE_LINK); }

// =====================================================================================================
// despeck_process: src/despeck/despeck.cpp:14-500 -- same checks, error codes, block schedule and output
// (ENVI, one band: Float32 for a single input band, CFloat32 for an interferogram / coherence).
int despeck_process(despeckOptions* opts) {
    opts->print();
    const int Nx = opts->Nx, Ny = opts->Ny;
    const int nulong = fringe_nulong(Nx, Ny);
    std::cout << "Number of uint32 bytes for mask: " << nulong << "\n";
    Raster in;
    if (!in.open(opts->inputDS)) {
        std::cout << "Cannot open stack file { " << opts->inputDS << " } for reading: " << in.error << "\nExiting with error code .... (102) \n";
        return 102;
    }
    if (!in.is_cfloat32_stack()) {
        std::cout << "Input stack { " << opts->inputDS << " } is not a VRT with one flat CFloat32 file per band\nExiting with error code .... (102) \n";
        return 102;
    }
    const int cols = in.cols, rows = in.rows, nbands = in.count();
    std::cout << "Number of rows  = " << rows << "\nNumber of cols  = " << cols << "\nNumber of bands = " << nbands << "\n";
    const int band1 = opts->ibands[0], band2 = opts->ibands[1];
    if (band1 <= 0 || band1 > nbands) { std::cout << "Master band " << band1 << " outside the range of permissible bands\nExiting with error code ... (102) \n"; return 102; }
    if (band2 > 0 && band2 > nbands) { std::cout << "Slave band " << band2 << "outside the range of permissible bands\nExiting with error code ... (102) \n"; return 102; }
    Raster wts;
    if (!wts.open(opts->wtsDS)) { std::cout << "Could not open weights dataset {" << opts->wtsDS << "}\nExiting with non-zero error code ... 105 \n"; return 105; }
    {
        int code = 0;
        if (wts.cols != cols) { std::cout << "Width mismatch between input dataset and weight dataset\n"; code = 106; }
        if (wts.rows != rows) { std::cout << "Length mismatch between input dataset and weight dataset \n"; code = 107; }
        if (wts.count() != nulong) { std::cout << "Number of bands mismatch for weights and window size \n"; code = 108; }
        auto geti = [&](const char* k) { auto it = wts.envi.fields.find(k); return it == wts.envi.fields.end() ? 0 : std::atoi(it->second.c_str()); };
        const int inNx = geti("halfwindowx"), inNy = geti("halfwindowy");
        if (inNx != Nx) { std::cout << "Half window size x of wts is different from input. \n"; code = 109; }
        if (inNx == 0) { std::cout << "No non-zero metadata item called HALFWINDOWX \n"; code = 109; }
        if (inNy != Ny) { std::cout << "Half window size y of wts is different from input. \n"; code = 110; }
        if (inNy == 0) { std::cout << "No non-zero metadata item called HALFWINDOWY \n"; code = 110; }
        if (code) { std::cout << "Exiting with error code ....(" << code << ")\n"; return code; }
        if (!wts.interleaved || wts.envi.data_type != 13) { std::cout << "Weights dataset must be a UInt32 ENVI BIP raster\n"; return 105; }
    }
    const int ngpu = visible_gpus();
    if (ngpu <= 0) { std::cout << "No CUDA device available and there is no CPU path.\n"; return 200 + FRINGE_ERR_NO_DEVICE; }

    const int blockysize = block_height(std::max(1, opts->memsize / workers_for(ngpu)), cols, opts->blocksize, 4 * (6 + nulong), rows, Ny);   // despeck.cpp:155
    std::cout << "Block size = " << blockysize << " lines \n";
    const std::vector<Block> sched = make_schedule(rows, blockysize, Ny);
    std::cout << "Total number of blocks to process: " << sched.size() << "\n";

    const bool cplx = band2 > 0;
    EnviWriter wout;
    if (!wout.create(opts->outputDS, cols, rows, 1, cplx ? 6 : 4)) {
        std::cout << "Could not create despecked dataset {" << opts->outputDS << "} \nExiting with non-zero error code ... 104 \n";
        return 104;
    }
    std::atomic<size_t> next(0);
    std::atomic<int> rc(0);
    std::mutex log_mu;
    auto worker = [&](int dev) {
        fringe_ctx* ctx = nullptr;
        if (fringe_create(dev, &ctx) != FRINGE_OK) { rc = 200 + FRINGE_ERR_NO_DEVICE; return; }
        const size_t bp = (size_t)cols * blockysize;
        Pinned z1, z2, wbuf, out, real;
        if (!z1.alloc(bp * 8) || !z2.alloc(bp * 8) || !wbuf.alloc(bp * nulong * 4) || !out.alloc(bp * 8) || !real.alloc(bp * 4)) {
            rc = 200 + FRINGE_ERR_MEMORY; fringe_destroy(ctx); return;
        }
        for (size_t i = next++; i < sched.size() && rc == 0; i = next++) {
            const Block& b = sched[i];
            bool ok = in.read_band_lines(band1 - 1, b.yoff, b.inysize, (char*)z1.p, 8);
            if (ok && cplx) ok = in.read_band_lines(band2 - 1, b.yoff, b.inysize, (char*)z2.p, 8);
            if (ok) ok = wts.read_interleaved_lines(b.yoff, b.inysize, wbuf.p);
            if (!ok) { std::lock_guard<std::mutex> g(log_mu); std::cout << "Error reading data at line " << b.yoff << "\nExiting with error code .... (108) \n"; rc = 108; break; }
            const int stt = fringe_despeck_block(ctx, (const float*)z1.p, cplx ? (const float*)z2.p : nullptr, (const uint32_t*)wbuf.p,
                                                 cols, b.inysize, Nx, Ny, b.first, b.nwrite, opts->computeCoherence ? 1 : 0, (float*)out.p);
            if (stt != FRINGE_OK) {
                std::lock_guard<std::mutex> g(log_mu);
                std::cout << "Device error: " << fringe_last_error(ctx) << "\n";
                rc = 200 + stt; break;
            }
            const size_t off = (size_t)b.first * cols, cnt = (size_t)b.nwrite * cols;
            if (cplx) ok = wout.write_lines(b.yoff + b.first, b.nwrite, (const char*)out.p + off * 8);
            else {                                   // Float32 output: real parts (GDAL converts the complex buffer)
                const float* src = (const float*)out.p + 2 * off;
                float* dst = (float*)real.p;
                for (size_t k = 0; k < cnt; ++k) dst[k] = src[2 * k];
                ok = wout.write_lines(b.yoff + b.first, b.nwrite, dst);
            }
            if (!ok) { std::lock_guard<std::mutex> g(log_mu); std::cout << "Error writing despeck data at line " << b.yoff << "\nExiting with error code .... (110) \n"; rc = 110; break; }
        }
        fringe_destroy(ctx);
    };
    {
        std::vector<std::thread> th;
        const int nw = (int)std::min<size_t>(workers_for(ngpu), sched.size());
        for (int d = 0; d < nw; ++d) th.emplace_back(worker, d % ngpu);
        for (auto& t : th) t.join();
    }
    if (rc != 0) return rc;
    wout.close_file();
    return 0;
}

// =====================================================================================================
// ampdispersion_process: src/ampdispersion/ampdispersion.cpp:14-330 -- calibration constants from the
// band metadata normalised by the reference band, non-overlapping blocks of lines, two Float32 ENVI
// rasters annotated with N = number of bands.
int ampdispersion_process(ampdispersionOptions* opts) {
    opts->print();
    Raster in;
    if (!in.open(opts->inputDS)) {
        std::cout << "Cannot open stack file { " << opts->inputDS << " } for reading: " << in.error << "\nExiting with error code .... (102) \n";
        return 102;
    }
    if (!in.is_cfloat32_stack()) {
        std::cout << "Input stack { " << opts->inputDS << " } is not a VRT with one flat CFloat32 file per band\nExiting with error code .... (102) \n";
        return 102;
    }
    const int cols = in.cols, rows = in.rows, nbands = in.count();
    std::cout << "Number of rows  = " << rows << "\nNumber of cols  = " << cols << "\nNumber of bands = " << nbands << "\n";
    // block height: the reference keeps one band of a block in memory at a time (ampdispersion.cpp:56), this
    // driver all of them (pinned, 8 bytes per band and pixel + the two outputs); blocks do not overlap, so
    // the height has no influence on the result
    int blockysize = int((opts->memsize * 1.0e6) / cols) / (opts->blocksize * (8 * nbands + 8)) * opts->blocksize;
    if (blockysize < opts->blocksize) blockysize = opts->blocksize;
    if (blockysize > rows) blockysize = rows;
    std::cout << "Block size = " << blockysize << " lines \n";
    std::cout << "Total number of blocks to process: " << (rows + blockysize - 1) / blockysize << "\n";

    std::vector<double> alpha(nbands, 1.0);
    for (int b = 0; b < nbands; ++b) {
        double c = 0.0;
        if (!in.bands.empty()) {
            auto it = in.bands[b].md_slc.find("amplitudeConstant");
            if (it != in.bands[b].md_slc.end()) c = std::atof(it->second.c_str());
        }
        if (c <= 0.0) { c = 1.0; std::cout << "No calibration constant found for band " << b + 1 << ". Setting to 1.0. \n"; }
        alpha[b] = c;
    }
    if (opts->refband < 1 || opts->refband > nbands) {
        std::cout << "Reference band number: " << opts->refband << " is invalid \nExiting with non-zero error code .... (102) \n";
        return 102;
    }
    {
        const double norm = alpha[opts->refband - 1];
        for (auto& a : alpha) a /= norm;
        alpha[opts->refband - 1] = 1.0;
    }
    for (int b = 0; b < nbands; ++b) std::cout << "Band " << b + 1 << ": " << alpha[b] << "\n";
    const int ngpu = visible_gpus();
    if (ngpu <= 0) { std::cout << "No CUDA device available and there is no CPU path.\n"; return 200 + FRINGE_ERR_NO_DEVICE; }

    EnviWriter wda, wmean;
    if (!wda.create(opts->daDS, cols, rows, 1, 4)) { std::cout << "Could not create ampdisp dataset {" << opts->daDS << "} \nExiting with non-zero error code ... 103 \n"; return 103; }
    if (!wmean.create(opts->meanampDS, cols, rows, 1, 4)) { std::cout << "Could not create meanamp dataset {" << opts->meanampDS << "} \nExiting with non-zero error code ... 104 \n"; return 104; }
    wda.set_metadata("N", std::to_string(nbands));
    wmean.set_metadata("N", std::to_string(nbands));

    std::vector<Block> sched;
    for (int yoff = 0; yoff < rows; yoff += blockysize) sched.push_back({yoff, std::min(blockysize, rows - yoff), 0, std::min(blockysize, rows - yoff)});
    std::atomic<size_t> next(0);
    std::atomic<int> rc(0);
    std::mutex log_mu;
    auto worker = [&](int dev) {
        fringe_ctx* ctx = nullptr;
        if (fringe_create(dev, &ctx) != FRINGE_OK) { rc = 200 + FRINGE_ERR_NO_DEVICE; return; }
        const size_t bp = (size_t)cols * blockysize;
        Pinned slc, da, mean;
        if (!slc.alloc(bp * nbands * 8) || !da.alloc(bp * 4) || !mean.alloc(bp * 4)) { rc = 200 + FRINGE_ERR_MEMORY; fringe_destroy(ctx); return; }
        for (size_t i = next++; i < sched.size() && rc == 0; i = next++) {
            const Block& b = sched[i];
            const size_t np = (size_t)cols * b.inysize;
            bool ok = true;
            for (int band = 0; band < nbands && ok; ++band)
                ok = in.read_band_lines(band, b.yoff, b.inysize, (char*)slc.p + (size_t)band * np * 8, 8);
            if (!ok) { std::lock_guard<std::mutex> g(log_mu); std::cout << "Error reading data at line " << b.yoff << "\nExiting with error code .... (108) \n"; rc = 108; break; }
            const int stt = fringe_ampdispersion_block(ctx, (const float*)slc.p, alpha.data(), cols, b.inysize, nbands, (float*)da.p, (float*)mean.p);
            if (stt != FRINGE_OK) {
                std::lock_guard<std::mutex> g(log_mu);
                std::cout << "Device error: " << fringe_last_error(ctx) << "\n";
                rc = 200 + stt; break;
            }
            if (!wda.write_lines(b.yoff, b.inysize, da.p)) { rc = 109; break; }
            if (!wmean.write_lines(b.yoff, b.inysize, mean.p)) { rc = 110; break; }
        }
        fringe_destroy(ctx);
    };
    {
        std::vector<std::thread> th;
        const int nw = (int)std::min<size_t>(workers_for(ngpu), sched.size());
        for (int d = 0; d < nw; ++d) th.emplace_back(worker, d % ngpu);
        for (auto& t : th) t.join();
    }
    if (rc != 0) return rc;
    wda.close_file();
    wmean.close_file();
    return 0;
}

// =====================================================================================================
// calamp_process: src/calamp/calamp.cpp:14-275 -- amplitude calibration constant of every band (mean amplitude of
// its valid pixels), written as <MDI key="amplitudeConstant"> into the slc metadata domain of a copy of the stack VRT
// (the producer side of nmap.cpp:204-233 / ampdispersion.cpp:100-127).  Same checks and error codes (102 open, 104-106
// mask size, 107 / 108 read errors); blocks do not overlap.
namespace {
// copy of the VRT text with amplitudeConstant set per band and relative source paths made absolute (the copy may live
// in another folder; GDAL's CreateCopy does the same)
std::string vrt_with_constants(const std::string& txt, const std::string& src_dir, const std::vector<std::string>& values) {
    std::string out;
    size_t pos = 0;
    int band = 0;
    while (true) {
        const size_t a = txt.find("<VRTRasterBand", pos);
        if (a == std::string::npos) { out += txt.substr(pos); break; }
        const size_t z = txt.find("</VRTRasterBand>", a);
        if (z == std::string::npos) { out += txt.substr(pos); break; }
        out += txt.substr(pos, a - pos);
        std::string body = txt.substr(a, z - a);
        // relative source paths -> absolute
        size_t q = 0;
        while ((q = body.find("relativeToVRT=\"1\"", q)) != std::string::npos) {
            const size_t gt = body.find('>', q), lt = body.find('<', gt);
            if (gt == std::string::npos || lt == std::string::npos) break;
            const std::string fn = fringe_host::trim(body.substr(gt + 1, lt - gt - 1));
            body.replace(gt + 1, lt - gt - 1, fringe_host::join_path(src_dir, fn));
            body.replace(q, 17, "relativeToVRT=\"0\"");
            q += 17;
        }
        const std::string mdi = "<MDI key=\"amplitudeConstant\">" + values[band] + "</MDI>";
        size_t m = body.find("<Metadata domain=\"slc\">");
        if (m != std::string::npos) {
            const size_t me = body.find("</Metadata>", m);
            // drop an existing constant
            const size_t old = body.find("<MDI key=\"amplitudeConstant\">", m);
            if (old != std::string::npos && old < me) {
                const size_t oe = body.find("</MDI>", old);
                body.erase(old, oe + 6 - old);
            }
            const size_t me2 = body.find("</Metadata>", m);
            body.insert(me2, "    " + mdi + "\n        ");
        } else {
            body += "    <Metadata domain=\"slc\">\n            " + mdi + "\n        </Metadata>\n    ";
        }
        out += body + "</VRTRasterBand>";
        pos = z + 16;
        ++band;
        if (band >= (int)values.size()) { out += txt.substr(pos); break; }
    }
    return out;
}
}  // namespace

int calamp_process(calampOptions* opts) {
    opts->print();
    Raster in;
    if (!in.open(opts->inputDS)) {
        std::cout << "Cannot open stack file { " << opts->inputDS << " } for reading: " << in.error << "\nExiting with error code .... (102) \n";
        return 102;
    }
    if (!in.is_cfloat32_stack()) {
        std::cout << "Input stack { " << opts->inputDS << " } is not a VRT with one flat CFloat32 file per band\nExiting with error code .... (102) \n";
        return 102;
    }
    const int cols = in.cols, rows = in.rows, nbands = in.count();
    std::cout << "Number of rows  = " << rows << "\nNumber of cols  = " << cols << "\nNumber of bands = " << nbands << "\n";
    Raster msk;
    bool have_mask = false;
    if (!opts->maskDS.empty()) {
        if (!msk.open(opts->maskDS)) { std::cout << "Cannot open mask file { " << opts->maskDS << " } for reading. \nExiting with error code .... (102) \n"; return 102; }
        int code = 0;
        if (msk.cols != cols) { std::cout << "Mask file width does not match stack size width \n"; code = 104; }
        if (msk.rows != rows) { std::cout << "Mask file length does not match stack size length \n"; code = 105; }
        if (msk.count() != 1) { std::cout << "Mask file has more than one band \n"; code = 106; }
        if (code) { std::cout << "Exiting with error code .... (" << code << ")\n"; return code; }
        have_mask = true;
    }
    const int ngpu = visible_gpus();
    if (ngpu <= 0) { std::cout << "No CUDA device available and there is no CPU path.\n"; return 200 + FRINGE_ERR_NO_DEVICE; }
    // the reference keeps one band of a block in memory (calamp.cpp:107), this driver all of them (pinned)
    int blockysize = opts->blocksize * (int((opts->memsize * 1.0e6) / cols) / (opts->blocksize * 8 * std::max(1, nbands)));
    if (blockysize < opts->blocksize) blockysize = opts->blocksize;
    if (blockysize > rows) blockysize = rows;
    std::cout << "Block size = " << blockysize << " lines \n";
    std::cout << "Total number of blocks to process: " << (rows + blockysize - 1) / blockysize << "\n";

    std::vector<double> totalsum(nbands, 0.0), norms(nbands, 0.0);
    fringe_ctx* ctx = nullptr;
    if (fringe_create(0, &ctx) != FRINGE_OK) return 200 + FRINGE_ERR_NO_DEVICE;
    const size_t bp = (size_t)cols * blockysize;
    Pinned slc, mask;
    std::vector<char> mask_scratch;
    if (!slc.alloc(bp * nbands * 8) || !mask.alloc(bp)) { fringe_destroy(ctx); return 200 + FRINGE_ERR_MEMORY; }
    int rc = 0;
    for (int yoff = 0; yoff < rows && rc == 0; yoff += blockysize) {
        const int inysize = std::min(blockysize, rows - yoff);
        const size_t np = (size_t)cols * inysize;
        if (have_mask && !msk.read_mask_lines(yoff, inysize, (uint8_t*)mask.p, mask_scratch)) {
            std::cout << "Error reading mask at line " << yoff << "\nExiting with error code .... (107) \n"; rc = 107; break;
        }
        for (int band = 0; band < nbands; ++band)
            if (!in.read_band_lines(band, yoff, inysize, (char*)slc.p + (size_t)band * np * 8, 8)) {
                std::cout << "Error reading data from band " << band + 1 << " at line " << yoff << "\nExiting with error code .... (108) \n"; rc = 108; break;
            }
        if (rc) break;
        const int st = fringe_calamp_block(ctx, (const float*)slc.p, have_mask ? (const uint8_t*)mask.p : nullptr, cols, inysize, nbands,
                                           totalsum.data(), norms.data());
        if (st != FRINGE_OK) { std::cout << "Device error: " << fringe_last_error(ctx) << "\n"; rc = 200 + st; }
    }
    fringe_destroy(ctx);
    if (rc) return rc;

    std::cout << "Normalization coefficients \n";
    std::vector<std::string> values(nbands);
    for (int b = 0; b < nbands; ++b) {
        if (norms[b] == 0) { std::cout << "No valid data found in Band " << b + 1 << ". Set to 1.0 \n"; totalsum[b] = 1.0; }
        else {
            totalsum[b] /= norms[b];
            std::cout << "Band " << b + 1 << ":" << totalsum[b] << "  from  " << int(10000 * norms[b] / (1.0 * rows * cols)) / 100.0 << " % of image \n";
        }
        std::ostringstream strs;                       // calamp.cpp:258-260: default stream formatting (6 significant digits)
        strs << totalsum[b];
        values[b] = strs.str();
    }
    std::string txt;
    if (!fringe_host::slurp(opts->inputDS, txt)) return 102;
    std::ofstream o(opts->outputDS.c_str());
    if (!o) { std::cout << "Could not create output VRT {" << opts->outputDS << "}\n"; return 103; }
    // dirname of the input, absolute
    std::string dir = fringe_host::dirname_of(opts->inputDS);
    if (dir.empty() || dir[0] != '/') { char buf[4096]; if (::getcwd(buf, sizeof(buf))) dir = std::string(buf) + "/" + dir; }
    o << vrt_with_constants(txt, dir, values);
    return o ? 0 : 103;
}
