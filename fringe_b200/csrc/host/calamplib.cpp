// Python module `calamplib` -- same class surface as the reference's Cython wrapper (src/calamp/calamplib.pyx): class
// Calamp with properties inputDS, outputDS, maskDS, defaultValue, applySqrt, blocksize, memsize and methods print(), run().
#include <pybind11/pybind11.h>

#include "options.hpp"

namespace py = pybind11;

PYBIND11_MODULE(calamplib, m) {
    m.doc() = "B200-native drop-in for FRInGE's calamplib";
    py::class_<calampOptions>(m, "Calamp", py::module_local())
        .def(py::init<>())
        .def_readwrite("inputDS", &calampOptions::inputDS)
        .def_readwrite("outputDS", &calampOptions::outputDS)
        .def_readwrite("maskDS", &calampOptions::maskDS)
        .def_readwrite("defaultValue", &calampOptions::defaultValue)
        .def_readwrite("applySqrt", &calampOptions::applySqrt)
        .def_readwrite("blocksize", &calampOptions::blocksize)
        .def_readwrite("memsize", &calampOptions::memsize)
        .def("print", [](calampOptions& self) { self.print(); })
        .def("run", [](calampOptions& self) {
            int rc;
            {
                py::gil_scoped_release nogil;
                rc = calamp_process(&self);
            }
            if (rc != 0) throw std::runtime_error("calamp_process returned " + std::to_string(rc));
        });
}
