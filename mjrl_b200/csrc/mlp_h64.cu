#include "mlp_inst.cuh"
namespace mjb {
MJB_DEFINE_MLP(64, 128)
}
