// Python module `ampdispersionlib` -- same class surface as the reference's Cython wrapper
// (src/ampdispersion/ampdispersionlib.pyx:20-90): class Ampdispersion with properties inputDS, meanampDS,
// outputDS, blocksize, memsize, refband and methods print(), run().
#include <pybind11/pybind11.h>

#include "options.hpp"

namespace py = pybind11;

PYBIND11_MODULE(ampdispersionlib, m) {
    m.doc() = "B200-native drop-in for FRInGE's ampdispersionlib";
    py::class_<ampdispersionOptions>(m, "Ampdispersion", py::module_local())
        .def(py::init<>())
        .def_readwrite("inputDS", &ampdispersionOptions::inputDS)
        .def_readwrite("meanampDS", &ampdispersionOptions::meanampDS)
        .def_readwrite("outputDS", &ampdispersionOptions::daDS)
        .def_readwrite("blocksize", &ampdispersionOptions::blocksize)
        .def_readwrite("memsize", &ampdispersionOptions::memsize)
        .def_readwrite("refband", &ampdispersionOptions::refband)
        .def("print", [](ampdispersionOptions& self) { self.print(); })
        .def("run", [](ampdispersionOptions& self) {
            int rc;
            {
                py::gil_scoped_release nogil;
                rc = ampdispersion_process(&self);
            }
            if (rc != 0) throw std::runtime_error("ampdispersion_process returned " + std::to_string(rc));
        });
}
