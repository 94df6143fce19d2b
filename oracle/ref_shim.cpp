// ref_shim.cpp -- extern "C" doorway onto the UNMODIFIED reference C++.
//
// TEST INFRASTRUCTURE ONLY (see oracle/perm_oracle.c header).  This file holds
// no reference code: it includes the reference's own headers and is compiled
// together with the reference's own src/permanent.cpp and
// src/permanent_laplace.cpp, where they lie under /root/reference, by
// oracle/build.py into oracle/_ref/libpqref.so (git-ignored).  It is what
// piquasso/_math/permanent.cpp:26-58 does, minus pybind11.
#include <complex>
#include <cstdint>
#include <cstring>
#include <string>

#include "matrix.hpp"
#include "n_aryGrayCodeCounter.hpp"
#include "permanent.hpp"
#include "permanent_laplace.hpp"

extern "C" {

// permanent_cpp<double>; returns 1 where the reference throws (sum mismatch).
int pqref_permanent_c128(const double *A, int R, int C, const int *rows,
                         const int *cols, double out[2])
{
    Matrix<std::complex<double>> m(
        (size_t)R, (size_t)C,
        reinterpret_cast<std::complex<double> *>(const_cast<double *>(A)));
    Vector<int> r((size_t)R, const_cast<int *>(rows));
    Vector<int> c((size_t)C, const_cast<int *>(cols));
    try {
        std::complex<double> v = permanent_cpp<double>(m, r, c);
        out[0] = v.real();
        out[1] = v.imag();
    } catch (std::string &) {
        return 1;
    }
    return 0;
}

int pqref_permanent_c64(const float *A, int R, int C, const int *rows,
                        const int *cols, float out[2])
{
    Matrix<std::complex<float>> m(
        (size_t)R, (size_t)C,
        reinterpret_cast<std::complex<float> *>(const_cast<float *>(A)));
    Vector<int> r((size_t)R, const_cast<int *>(rows));
    Vector<int> c((size_t)C, const_cast<int *>(cols));
    try {
        std::complex<float> v = permanent_cpp<float>(m, r, c);
        out[0] = v.real();
        out[1] = v.imag();
    } catch (std::string &) {
        return 1;
    }
    return 0;
}

// permanent_laplace_cpp<double>; out must hold 2*max(C,1) doubles.
int pqref_permanent_laplace_c128(const double *A, int R, int C, const int *rows,
                                 const int *cols, double *out, int *out_len)
{
    Matrix<std::complex<double>> m(
        (size_t)R, (size_t)C,
        reinterpret_cast<std::complex<double> *>(const_cast<double *>(A)));
    Vector<int> r((size_t)R, const_cast<int *>(rows));
    Vector<int> c((size_t)C, const_cast<int *>(cols));
    Vector<std::complex<double>> v = permanent_laplace_cpp<double>(m, r, c);
    *out_len = (int)v.size();
    for (size_t i = 0; i < v.size(); i++) {
        out[2 * i] = v[i].real();
        out[2 * i + 1] = v[i].imag();
    }
    return 0;
}

int pqref_permanent_laplace_c64(const float *A, int R, int C, const int *rows,
                                const int *cols, float *out, int *out_len)
{
    Matrix<std::complex<float>> m(
        (size_t)R, (size_t)C,
        reinterpret_cast<std::complex<float> *>(const_cast<float *>(A)));
    Vector<int> r((size_t)R, const_cast<int *>(rows));
    Vector<int> c((size_t)C, const_cast<int *>(cols));
    Vector<std::complex<float>> v = permanent_laplace_cpp<float>(m, r, c);
    *out_len = (int)v.size();
    for (size_t i = 0; i < v.size(); i++) {
        out[2 * i] = v[i].real();
        out[2 * i + 1] = v[i].imag();
    }
    return 0;
}

// The reference counter itself: Gray digits at `offset`, then `nsteps` calls of
// next(); trace receives (changed, prev, value) triples.  Valid for
// offset < 2^31 only (reference int truncation, n_aryGrayCodeCounter.hpp:179).
int pqref_gray_trace(const int *limits, int ndigits, int64_t offset, int nsteps,
                     int *gray0, int *trace)
{
    n_aryGrayCodeCounter counter(const_cast<int *>(limits), (size_t)ndigits, offset);
    std::memcpy(gray0, counter.get(), sizeof(int) * (size_t)ndigits);
    int done = 0;
    for (int s = 0; s < nsteps; s++) {
        int changed = 0, prev = 0, value = 0;
        if (counter.next(changed, prev, value))
            break;
        trace[3 * s] = changed;
        trace[3 * s + 1] = prev;
        trace[3 * s + 2] = value;
        done++;
    }
    return done;
}

} // extern "C"
