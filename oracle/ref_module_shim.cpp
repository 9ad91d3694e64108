// Test infrastructure: pybind11 module definition for the UNMODIFIED reference C++
// residuals (compiled in place from /root/reference/c++/residual_function.cpp and
// residual_function_omp.cpp by oracle/Makefile; outputs only under oracle/_ref/).
// The reference's own module file (c++/_residual_function.cpp:9-14) uses
// PYBIND11_PLUGIN, which pybind11 >= 3 no longer provides; this shim exports the
// same module name and the same two functions.
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>

namespace py = pybind11;

py::array residual_function(py::array X, py::object graph);
py::array residual_function_omp(py::array X, py::object graph);

PYBIND11_MODULE(_residual_function, m)
{
    m.def("residual_function", &residual_function);
    m.def("residual_function_omp", &residual_function_omp);
}
