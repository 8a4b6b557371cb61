// kernels.cuh -- the sm_100a kernels of the ShippingEnv hot path, one header per group:
//   pipeline.cuh          constants + PTX helpers (mbarrier, TMA bulk copies, world staging)
//   step_kernels.cuh      general step kernel (all outputs, any action)
//   step_fast_kernels.cuh throughput step kernels (the roofline kernel of bench.py)
//   step_many_kernels.cuh T steps per launch with the ship in registers (shipenv_step_many)
//   aux_kernels.cuh       reset / observe / sample_action / valid-action mask / discretize
//   mcts_kernels.cuh      rollouts and the full MCTS choose_action
//
// Layout in HBM (DESIGN.md "Data layout"): ships[n] 16-byte records, clock[n] u32, actions[n] i16,
// reward[n] f32/f64, done/reset/error ballots ceil(n/32) u32 each.  One env per thread; every CTA is
// persistent, stages the world blob into shared memory with one TMA bulk copy and is fed 256-env tiles
// through a shared-memory ring filled by a TMA producer warp.
#pragma once
#include "pipeline.cuh"
#include "step_kernels.cuh"
#include "step_fast_kernels.cuh"
#include "step_many_kernels.cuh"
#include "aux_kernels.cuh"
#include "mcts_kernels.cuh"
