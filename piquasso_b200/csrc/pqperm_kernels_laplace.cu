// placeholder until the Laplace walk lands
#include <string>
#include "../../include/pqperm.h"
namespace pqperm {
int laplace_run(int, const double *, int, int, const int *, const int *, double *, int *,
                std::string &err)
{
    err = "permanent_laplace kernels not built";
    return PQ_ERR_CUDA;
}
} // namespace pqperm
extern "C" int pq_perm_laplace_batch_c128(int, const double *, const int64_t *, const int32_t *,
                                          const int32_t *, const int32_t *, const int64_t *,
                                          const int32_t *, const int64_t *, double *,
                                          const int64_t *, int32_t *)
{
    return PQ_ERR_CUDA;
}
