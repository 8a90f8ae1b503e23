// Python module `nmaplib` -- same class surface as the reference's Cython wrapper
// (src/nmap/nmaplib.pyx:27-135): class Nmap with properties inputDS, countDS, weightsDS, maskDS,
// method, minimumProbability, blocksize, memsize, halfWindowX, halfWindowY, noGPU and methods
// print(), run().  run() releases the GIL like the reference (`with nogil`, nmaplib.pyx:134);
// unlike the reference it does not drop the driver's return code: non-zero raises RuntimeError.
#include <pybind11/pybind11.h>

#include "options.hpp"

namespace py = pybind11;

PYBIND11_MODULE(nmaplib, m) {
    m.doc() = "B200-native drop-in for FRInGE's nmaplib";
    py::class_<nmapOptions>(m, "Nmap", py::module_local())
        .def(py::init<>())
        .def_readwrite("inputDS", &nmapOptions::inputDS)
        .def_readwrite("countDS", &nmapOptions::ncountDS)
        .def_readwrite("weightsDS", &nmapOptions::wtsDS)
        .def_readwrite("maskDS", &nmapOptions::maskDS)
        .def_readwrite("method", &nmapOptions::method)
        .def_readwrite("minimumProbability", &nmapOptions::prob)
        .def_readwrite("blocksize", &nmapOptions::blocksize)
        .def_readwrite("memsize", &nmapOptions::memsize)
        .def_readwrite("halfWindowX", &nmapOptions::Nx)
        .def_readwrite("halfWindowY", &nmapOptions::Ny)
        .def_readwrite("noGPU", &nmapOptions::noGPU)
        .def("print", [](nmapOptions& self) { self.print(); })
        .def("run", [](nmapOptions& self) {
            int rc;
            {
                py::gil_scoped_release nogil;
                rc = nmap_process(&self);
            }
            if (rc != 0) throw std::runtime_error("nmap_process returned " + std::to_string(rc));
        });
}
