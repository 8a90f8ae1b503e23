// Python module `despecklib` -- same class surface as the reference's Cython wrapper
// (src/despeck/despecklib.pyx): class Despeck with properties inputDS, weightsDS, outputDS, blocksize,
// memsize, halfWindowX, halfWindowY, band1, band2, coherenceFlag and methods print(), run().
// run() releases the GIL; a non-zero driver return code raises RuntimeError.
#include <pybind11/pybind11.h>

#include "options.hpp"

namespace py = pybind11;

PYBIND11_MODULE(despecklib, m) {
    m.doc() = "B200-native drop-in for FRInGE's despecklib";
    py::class_<despeckOptions>(m, "Despeck", py::module_local())
        .def(py::init<>())
        .def_readwrite("inputDS", &despeckOptions::inputDS)
        .def_readwrite("weightsDS", &despeckOptions::wtsDS)
        .def_readwrite("outputDS", &despeckOptions::outputDS)
        .def_readwrite("blocksize", &despeckOptions::blocksize)
        .def_readwrite("memsize", &despeckOptions::memsize)
        .def_readwrite("halfWindowX", &despeckOptions::Nx)
        .def_readwrite("halfWindowY", &despeckOptions::Ny)
        .def_property("band1", [](despeckOptions& s) { return s.ibands[0]; }, [](despeckOptions& s, int v) { s.ibands[0] = v; })
        .def_property("band2", [](despeckOptions& s) { return s.ibands[1]; }, [](despeckOptions& s, int v) { s.ibands[1] = v; })
        .def_readwrite("coherenceFlag", &despeckOptions::computeCoherence)
        .def("print", [](despeckOptions& self) { self.print(); })
        .def("run", [](despeckOptions& self) {
            int rc;
            {
                py::gil_scoped_release nogil;
                rc = despeck_process(&self);
            }
            if (rc != 0) throw std::runtime_error("despeck_process returned " + std::to_string(rc));
        });
}
