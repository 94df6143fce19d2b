// pybind_permanent.cpp -- the drop-in for the reference's native module
// `piquasso._math.permanent` (piquasso/_math/permanent.cpp:26-85 of the
// reference): same module name, same four overloads in the same order, same
// argument names, same return types (0-d / 1-d numpy arrays of the matrix's
// complex dtype).  The arithmetic is one call into libpqperm.so's C ABI
// (include/pqperm.h); the GIL is released around it.
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <complex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pqperm.h"

namespace py = pybind11;

namespace {

// array_t<int, c_style | forcecast>: lists, tuples and int64 arrays are accepted
using IntArray = py::array_t<int, py::array::c_style | py::array::forcecast>;

[[noreturn]] void raise_for(int rc)
{
    const std::string msg = pq_last_error();
    if (rc == PQ_ERR_BAD_ARG || rc == PQ_ERR_TOO_LARGE)
        throw py::value_error(msg);
    // PQ_ERR_SUM_MISMATCH: the reference throws std::string, which pybind11
    // turns into RuntimeError
    throw std::runtime_error(msg);
}

void check_shapes(const py::buffer_info &m, const py::buffer_info &r, const py::buffer_info &c)
{
    if (m.ndim != 2)
        throw py::value_error("matrix must be 2-dimensional");
    if (r.ndim != 1 || c.ndim != 1)
        throw py::value_error("rows and cols must be 1-dimensional");
    if (r.shape[0] != m.shape[0] || c.shape[0] != m.shape[1])
        throw py::value_error("multiplicity lengths do not match the matrix shape");
}

template <typename T> struct Abi;
template <> struct Abi<double> {
    static int perm(const double *A, int R, int C, const int *r, const int *c, double *o)
    {
        return pq_perm_c128(A, R, C, r, c, o);
    }
    static int laplace(const double *A, int R, int C, const int *r, const int *c, double *o,
                       int *n)
    {
        return pq_perm_laplace_c128(A, R, C, r, c, o, n);
    }
};
template <> struct Abi<float> {
    static int perm(const float *A, int R, int C, const int *r, const int *c, float *o)
    {
        return pq_perm_c64(A, R, C, r, c, o);
    }
    static int laplace(const float *A, int R, int C, const int *r, const int *c, float *o,
                       int *n)
    {
        return pq_perm_laplace_c64(A, R, C, r, c, o, n);
    }
};

template <typename T>
py::object permanent_np(py::array_t<std::complex<T>, py::array::c_style> matrix,
                        IntArray row_mult_arr, IntArray col_mult_arr)
{
    py::buffer_info m = matrix.request(), r = row_mult_arr.request(), c = col_mult_arr.request();
    check_shapes(m, r, c);
    T out[2] = {0, 0};
    int rc;
    {
        py::gil_scoped_release release;
        rc = Abi<T>::perm(static_cast<const T *>(m.ptr), (int)m.shape[0], (int)m.shape[1],
                          static_cast<const int *>(r.ptr), static_cast<const int *>(c.ptr), out);
    }
    if (rc != PQ_OK)
        raise_for(rc);
    // 0-d array, like create_numpy_scalar (src/numpy_utils.hpp:36-49)
    py::array_t<std::complex<T>> result(std::vector<py::ssize_t>{});
    *result.mutable_data() = std::complex<T>(out[0], out[1]);
    return std::move(result);
}

template <typename T>
py::object permanent_laplace_np(py::array_t<std::complex<T>, py::array::c_style> matrix,
                                IntArray row_mult_arr, IntArray col_mult_arr)
{
    py::buffer_info m = matrix.request(), r = row_mult_arr.request(), c = col_mult_arr.request();
    check_shapes(m, r, c);
    const py::ssize_t width = m.shape[1] > 0 ? m.shape[1] : 1;
    std::vector<T> out(2 * (size_t)width);
    int n = 0, rc;
    {
        py::gil_scoped_release release;
        rc = Abi<T>::laplace(static_cast<const T *>(m.ptr), (int)m.shape[0], (int)m.shape[1],
                             static_cast<const int *>(r.ptr), static_cast<const int *>(c.ptr),
                             out.data(), &n);
    }
    if (rc != PQ_OK)
        raise_for(rc);
    py::array_t<std::complex<T>> result((py::ssize_t)n);
    auto *dst = result.mutable_data();
    for (int i = 0; i < n; i++)
        dst[i] = std::complex<T>(out[2 * i], out[2 * i + 1]);
    return std::move(result);
}

const char *permanent_docstring = R"(
Calculates the permanent of a matrix, based on Eq. (8) of
https://arxiv.org/abs/2309.07027.  B200 (sm_100a) implementation.
)";

const char *permanent_laplace_docstring = R"(
Calculates the permanents of the submatrices corresponding to the Laplace expansion,
corresponding to Eq. (8) of https://arxiv.org/abs/2309.07027 and
Lemma 1 of https://arxiv.org/abs/2005.04214.  B200 (sm_100a) implementation.
)";

} // namespace

PYBIND11_MODULE(permanent, m)
{
    // float overloads first, as in the reference (piquasso/_math/permanent.cpp:73-84)
    m.def("permanent", &permanent_np<float>, permanent_docstring, py::arg("matrix"),
          py::arg("rows"), py::arg("cols"));
    m.def("permanent", &permanent_np<double>, permanent_docstring, py::arg("matrix"),
          py::arg("rows"), py::arg("cols"));
    m.def("permanent_laplace", &permanent_laplace_np<float>, permanent_laplace_docstring,
          py::arg("matrix"), py::arg("rows"), py::arg("cols"));
    m.def("permanent_laplace", &permanent_laplace_np<double>, permanent_laplace_docstring,
          py::arg("matrix"), py::arg("rows"), py::arg("cols"));
}
