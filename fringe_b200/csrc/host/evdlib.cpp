// Python modules `evdlib` (class Evd, src/evd/evdlib.cpp:16-40) and `phase_linklib` (class
// Phaselink, src/phase_link/phase_linklib.pyx:46-157): identical attribute names.  One source,
// compiled twice (-DFRINGE_PHASE_LINK selects the module name, class name and driver).
#include <pybind11/pybind11.h>

#include "options.hpp"

namespace py = pybind11;

#ifdef FRINGE_PHASE_LINK
#define MODNAME phase_linklib
#define CLSNAME "Phaselink"
#define DRIVER phase_link_process
#else
#define MODNAME evdlib
#define CLSNAME "Evd"
#define DRIVER evd_process
#endif

PYBIND11_MODULE(MODNAME, m) {
    py::class_<evdOptions>(m, CLSNAME, py::module_local())
        .def(py::init<>())
        .def_readwrite("inputDS", &evdOptions::inputDS)
        .def_readwrite("outputFolder", &evdOptions::outputFolder)
        .def_readwrite("outputCompressedSlcFolder", &evdOptions::outputCompressedSlcFolder)
        .def_readwrite("compSlc", &evdOptions::compSlc)
        .def_readwrite("weightsDS", &evdOptions::wtsDS)
        .def_readwrite("minimumNeighbors", &evdOptions::minNeighbors)
        .def_readwrite("miniStackCount", &evdOptions::miniStackCount)
        .def_readwrite("blocksize", &evdOptions::blocksize)
        .def_readwrite("memsize", &evdOptions::memsize)
        .def_readwrite("halfWindowX", &evdOptions::Nx)
        .def_readwrite("halfWindowY", &evdOptions::Ny)
        .def_readwrite("method", &evdOptions::method)
        .def_readwrite("bandWidth", &evdOptions::bandWidth)
        .def("print", [](evdOptions& self) { self.print(); })
        .def("run", [](evdOptions& self) {
            int rc;
            {
                py::gil_scoped_release nogil;      // the reference's pybind path holds the GIL (evdlib.cpp:37-39)
                rc = DRIVER(&self);
            }
            if (rc != 0) throw std::runtime_error(std::string(CLSNAME) + " driver returned " + std::to_string(rc));
        });
}
