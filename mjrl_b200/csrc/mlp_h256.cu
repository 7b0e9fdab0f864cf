#include "mlp_inst.cuh"
namespace mjb {
MJB_DEFINE_MLP(256, 64)
}
