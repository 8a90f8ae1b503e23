// Option structs of the block drivers: the C++ side of the drop-in boundary.  Member names,
// types and defaults follow src/nmap/nmap.hpp:23-65, src/evd/evd.hpp:19-68 and
// src/phase_link/phase_link.hpp (the args.hxx command-line parsers are not reproduced: the
// Python bindings never call them).
#pragma once
#include <iostream>
#include <string>

struct nmapOptions {
    std::string inputDS;     // input VRT with SLCs as bands
    std::string maskDS;      // optional byte mask
    bool noGPU;              // kept for interface compatibility; there is no CPU path here
    std::string wtsDS;       // output neighbourhood bit mask
    std::string ncountDS;    // output neighbour count
    std::string method;      // KS2 or AD2
    int blocksize;           // block quantum in lines
    int memsize;             // MB the block buffers may use
    int Nx, Ny;              // half window sizes
    double prob;             // minimum p-value

    nmapOptions() : noGPU(false), method("KS2"), blocksize(64), memsize(512), Nx(5), Ny(5), prob(0.05) {}
    void print() const {
        std::cout << "Input Dataset: " << inputDS << std::endl;
        std::cout << "Weights Dataset: " << wtsDS << std::endl;
        std::cout << "Count Dataset: " << ncountDS << std::endl;
        std::cout << "Mask Dataset: " << maskDS << std::endl;
        std::cout << "Window size: " << Nx << " " << Ny << std::endl;
        std::cout << "Threshold : " << prob << std::endl;
        std::cout << "Memsize: " << memsize << " Mb \n";
        std::cout << "Blocksize: " << blocksize << " lines \n";
        std::cout << "GPU user request: " << !(noGPU) << " \n";
    }
};

struct evdOptions {
    std::string inputDS;                    // input VRT with SLCs as bands
    std::string wtsDS;                      // neighbourhood bit mask from nmap
    std::string compSlc;                    // name of the compressed SLC
    std::string outputCompressedSlcFolder;  // folder of the compressed SLC
    std::string outputFolder;               // output stack folder (must not exist)
    std::string coherence;                  // unused by the reference driver
    std::string compSLC;                    // written by the reference's CLI parser only
    int blocksize, memsize;
    int Nx, Ny;
    int minNeighbors;
    std::string method;                     // MLE / EVD / STBAS
    int miniStackCount;
    int bandWidth;

    evdOptions() : blocksize(64), memsize(2048), Nx(5), Ny(5), minNeighbors(2), method("MLE"),
                   miniStackCount(1), bandWidth(-1) {}
    void print() const {
        std::cout << "Input Dataset: " << inputDS << std::endl;
        std::cout << "Weights Dataset: " << wtsDS << std::endl;
        std::cout << "Output Folder " << outputFolder << std::endl;
        std::cout << "Output compressed SLC folder " << outputCompressedSlcFolder << std::endl;
        std::cout << "Output compressed SLC " << compSlc << std::endl;
        std::cout << "Window size: " << Nx << " " << Ny << std::endl;
        std::cout << "Memsize: " << memsize << " Mb \n";
        std::cout << "Blocksize: " << blocksize << " lines \n";
        std::cout << "Minimum neighbors: " << minNeighbors << "\n";
        std::cout << "Mini-stack counter: " << miniStackCount << "\n";
        std::cout << "Decomposition method: " << method << "\n";
        if (method.compare("STBAS") == 0) std::cout << "STBAS Bandwidth: " << bandWidth << "\n";
    }
};

// src/despeck/despeck.hpp:19-56
struct despeckOptions {
    std::string inputDS;     // input VRT with SLCs as bands
    std::string wtsDS;       // neighbourhood bit mask from nmap
    std::string outputDS;    // despeckled amplitude / interferogram / coherence
    int blocksize, memsize;
    int ibands[2];           // 1-based bands: master, slave (-1 = single-band amplitude)
    int Nx, Ny;
    bool computeCoherence;

    despeckOptions() : blocksize(64), memsize(512), Nx(5), Ny(5), computeCoherence(false) { ibands[0] = 1; ibands[1] = -1; }
    void print() const {
        std::cout << "Input Dataset: " << inputDS << std::endl;
        std::cout << "Weights Dataset: " << wtsDS << std::endl;
        std::cout << "Output Dataset: " << outputDS << std::endl;
        std::cout << "Bands: " << ibands[0] << " " << ibands[1] << std::endl;
        std::cout << "Window size: " << Nx << " " << Ny << std::endl;
        std::cout << "Memsize: " << memsize << " Mb \n";
        std::cout << "Blocksize: " << blocksize << " lines \n";
        std::cout << "Coherence: " << computeCoherence << " \n";
    }
};

// src/ampdispersion/ampdispersion.hpp:19-46
struct ampdispersionOptions {
    std::string inputDS;     // input VRT with SLCs as bands
    std::string meanampDS;   // output mean normalised amplitude
    std::string daDS;        // output amplitude dispersion
    int blocksize, memsize;
    int refband;             // 1-based reference band of the calibration constants

    ampdispersionOptions() : blocksize(64), memsize(256), refband(1) {}
    void print() const {
        std::cout << "Input Dataset: " << inputDS << std::endl;
        std::cout << "Dispersion Dataset: " << daDS << std::endl;
        std::cout << "Mean amplitude Dataset: " << meanampDS << std::endl;
        std::cout << "Memsize: " << memsize << " Mb \n";
        std::cout << "Blocksize: " << blocksize << " lines \n";
        std::cout << "Reference band: " << refband << " \n";
    }
};

// src/calamp/calamp.hpp:27-56
struct calampOptions {
    std::string inputDS;     // input VRT with SLCs as bands
    std::string maskDS;      // optional mask raster
    std::string outputDS;    // output VRT: copy of the input with amplitudeConstant in every band's slc metadata
    double defaultValue;     // carried like the reference's (calamp.cpp never reads it: bands without valid data get 1.0)
    bool applySqrt;          // likewise unused by the reference's loop
    int blocksize, memsize;

    calampOptions() : defaultValue(1.0), applySqrt(false), blocksize(128), memsize(256) {}
    void print() const {
        std::cout << "Input Dataset: " << inputDS << std::endl;
        std::cout << "Output Dataset: " << outputDS << std::endl;
        if (!maskDS.empty()) std::cout << "Mask Dataset: " << maskDS << std::endl;
        std::cout << "Memsize: " << memsize << " Mb \n";
        std::cout << "Blocksize: " << blocksize << " lines \n";
        std::cout << "Default norm: " << defaultValue << " \n";
    }
};

// Block drivers (drivers.cpp).  Return 0 or the reference's error codes
// (nmap: 1,102,104,105,106,108,111; evd: 101,102,105-110,112-121), plus 200+status when the
// device library reports an error (there is no CPU fallback).
int nmap_process(nmapOptions* opts);
int evd_process(evdOptions* opts);            // src/evd/evd.cpp control flow
int phase_link_process(evdOptions* opts);     // src/phase_link/phase_link.cpp control flow
int despeck_process(despeckOptions* opts);    // src/despeck/despeck.cpp: 102, 104-110, 200+status
int ampdispersion_process(ampdispersionOptions* opts);   // src/ampdispersion/ampdispersion.cpp: 102-104, 108-110
int calamp_process(calampOptions* opts);                 // src/calamp/calamp.cpp: 102-108
