// pgo.h -- pose-graph optimisation of lama::SimplePGO on the device (BASELINE config 5).
//
// Reference: SimplePGO::optimize src/simple_pgo.cpp:48-105 -- a prior on node 0 (sigmas 1) or on the fixed nodes (sigmas 0.1),
// BetweenFactor<SE2> on consecutive nodes (measured = node[i]^-1 node[i+1]) and on the loop edges, all with sigmas (0.5, 0.5, 0.1),
// optimised by miniSAM's Levenberg-Marquardt (vendor/minisam/minisam/nonlinear/LevenbergMarquardtOptimizer.cpp:56-332, defaults .h:21-36,
// outer loop nonlinear/NonlinearOptimizer.cpp:109-238).  Factor arithmetic: slam/BetweenFactor.h:50-67, slam/PriorFactor.h:52-64,
// geometry/Sophus.h:45-74 (Local = log(origin^-1 t), Retract = origin exp(v), Jacobians -Adj / Adj(v2^-1) / I), whitening
// core/LossFunction.cpp:95-114, normal equations nonlinear/linearization.cpp:150-341 (b = -J^T r).
//
// What is different on the device: the damped normal equations (J^T J + lambda diag(J^T J)) dx = b are solved by a block-Jacobi
// preconditioned conjugate gradient running in ONE cooperative kernel (grid-wide barriers between the phases of an iteration, all
// reductions in a fixed order) instead of Eigen's SimplicialLDLT: the system is symmetric positive definite, so the solution is the
// same up to the CG tolerance (relative residual 1e-10), and the LM decisions (gain ratio, lambda schedule, stop rule) follow the
// reference line by line on the host in fp64.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "lama_core.h"

namespace lama_b200 {

struct PgoEdge {
    int from, to;
    SE2 measured;
};
struct PgoFixed {
    int node;
    SE2 pose;
};
struct PgoReport {
    int status = -1;            // NonlinearOptimizationStatus: 0 SUCCESS, 1 MAX_ITERATION, 2 ERROR_INCREASE, 3 RANK_DEFICIENCY, 4 INVALID
    uint32_t iterations = 0;    // LM iterations (NonlinearOptimizer::iterations_)
    uint32_t lambda_tries = 0;  // tryLambda_ calls
    uint64_t cg_iterations = 0;
    double initial_error = 0, final_error = 0;
    double device_ms = 0;       // CUDA-event time of the whole optimisation
};

// SimplePGO::optimize: on SUCCESS `nodes` holds the optimised poses (like the reference, they are left untouched otherwise).
// Returns a LAMA_* status (0 = the call worked; the optimiser's verdict is in report.status).
int pgo_optimize(int device, std::vector<SE2>& nodes, const std::vector<PgoEdge>& edges, const std::vector<PgoFixed>& fixed, PgoReport& report, std::string& err);

}  // namespace lama_b200
