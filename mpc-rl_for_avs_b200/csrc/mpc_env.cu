// one step of the batched synthetic intersection environment as ONE kernel
// (one thread per environment).  Same rules and the same counter-based random numbers as the tensor
// program in rl.BatchedIntersectionEnv.step (which stays the CPU / reference implementation and
// the specification; rl.py cites the reference files the rules come from).  Not part of the MPC
// hot path: it exists so that BASELINE config 4 (RL in the loop) is not bound by ~250 tiny launches
// per environment step.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/mpc_b200.h"

namespace {

constexpr int kMaxOthers = 15;
constexpr uint64_t kCtrStep = 0xD1B54A32D192ED03ull;
constexpr float kPiF = 3.14159265358979323846f;

// splitmix64 of (draw counter, env * 64 + column) -> uniform in (0, 1); rl.BatchedIntersectionEnv._rand
__device__ __forceinline__ float u01(uint64_t ctr, int env, int col) {
  uint64_t x = (uint64_t)((int64_t)env * 64 + col) * 0x9E3779B97F4A7C15ull + ctr;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x = x ^ (x >> 31);
  return ((float)((x >> 40) & 0xFFFFFFull) + 0.5f) * (1.0f / 16777216.0f);
}
// _randn(n): Box-Muller on columns (i, n + i) of one _rand(2 n) call
__device__ __forceinline__ float gauss(uint64_t ctr, int env, int i, int n) {
  return sqrtf(-2.0f * logf(u01(ctr, env, i))) * cosf(2.0f * kPiF * u01(ctr, env, n + i));
}

struct Veh { float x, y, speed, heading; };

// _spawn_others(): four draw calls starting after `ctr` (ctr + 1..4 steps)
__device__ __forceinline__ Veh spawn(uint64_t ctr, int env, int m, int M) {
  const float lane_heading[4] = {-kPiF / 2, 0.0f, kPiF / 2, kPiF};
  float c = floorf(u01(ctr + 1 * kCtrStep, env, m) * 4.0f);
  c = c > 3.0f ? 3.0f : c;
  const float lane = 2.0f + 0.2f * gauss(ctr + 2 * kCtrStep, env, m, M);
  const float d = -30.0f + 100.0f * u01(ctr + 3 * kCtrStep, env, m);
  const float ang = c * (kPiF / 2);
  float spd = 8.0f + gauss(ctr + 4 * kCtrStep, env, m, M);
  spd = spd < 0.5f ? 0.5f : spd;
  Veh v;
  v.x = cosf(ang) * lane - sinf(ang) * d;
  v.y = sinf(ang) * lane + cosf(ang) * d;
  v.speed = spd;
  v.heading = lane_heading[(int)c];
  return v;
}

// observe(): ego row, then the others sorted by distance to the ego
__device__ void write_obs(float* o, int V, const float* ego, const Veh* oth, int M) {
  int order[kMaxOthers];
  float dist[kMaxOthers];
  for (int m = 0; m < M; ++m) {
    const float dx = oth[m].x - ego[0], dy = oth[m].y - ego[1];
    const float d = hypotf(dx, dy);
    int j = m;
    while (j > 0 && dist[j - 1] > d) { dist[j] = dist[j - 1]; order[j] = order[j - 1]; --j; }
    dist[j] = d; order[j] = m;
  }
  float s, c;
  sincosf(ego[2], &s, &c);
  o[0] = 1.0f; o[1] = ego[0]; o[2] = ego[1]; o[3] = ego[3] * c; o[4] = ego[3] * s; o[5] = ego[2]; o[6] = s; o[7] = c;
  for (int r = 0; r < M && r + 1 < V; ++r) {
    const Veh& v = oth[order[r]];
    sincosf(v.heading, &s, &c);
    float* q = o + (size_t)(r + 1) * 8;
    q[0] = 1.0f; q[1] = v.x; q[2] = v.y; q[3] = v.speed * c; q[4] = v.speed * s; q[5] = v.heading; q[6] = s; q[7] = c;
  }
}

__global__ void __launch_bounds__(128)
k_env_step(const MpcEnvStep a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const int M = a.n_others, V = M + 1;
  const uint64_t ctr = (uint64_t)(*a.counter);
  // ---- action -> physical controls (agents/a2c_mpc.py:151-153 -> ContinuousAction clip + lmap, quirk Q8)
  float acc = a.action[2 * b], steer = a.action[2 * b + 1];
  if (a.raw_action) {
    acc = fminf(fmaxf(acc, -1.0f), 1.0f) * 5.0f;
    steer = fminf(fmaxf(steer, -1.0f), 1.0f) * (kPiF / 4);
  } else {
    acc = fminf(fmaxf(acc, -5.0f), 5.0f);
    steer = fminf(fmaxf(steer, -kPiF / 4), kPiF / 4);
  }
  float ego[4] = {a.ego[4 * b], a.ego[4 * b + 1], a.ego[4 * b + 2], a.ego[4 * b + 3]};
  const float beta = atanf(0.5f * tanf(steer));
  for (int i = 0; i < a.substeps; ++i) {                    // highway-env Vehicle.step at the simulation frequency
    const float x = ego[0], y = ego[1], th = ego[2], v = ego[3];
    ego[0] = x + v * cosf(th + beta) * a.dt_sim;
    ego[1] = y + v * sinf(th + beta) * a.dt_sim;
    ego[2] = th + v * sinf(beta) / 2.5f * a.dt_sim;
    ego[3] = fminf(fmaxf(v + acc * a.dt_sim, -40.0f), 40.0f);
  }
  ego[2] = atan2f(sinf(ego[2]), cosf(ego[2]));
  // ---- others: constant velocity; those that left the map re-enter at its edge (draw calls 1-4)
  Veh oth[kMaxOthers];
  bool crashed = false;
  for (int m = 0; m < M; ++m) {
    const float* q = a.others + ((size_t)b * M + m) * 4;
    Veh v{q[0], q[1], q[2], q[3]};
    const float step_len = v.speed * ((float)a.substeps * a.dt_sim);
    v.x += step_len * cosf(v.heading);
    v.y += step_len * sinf(v.heading);
    if (fabsf(v.x) > 90.0f || fabsf(v.y) > 90.0f) {
      Veh f = spawn(ctr, b, m, M);
      if (f.heading == 0.0f) f.x = -80.0f; else if (f.heading == kPiF) f.x = 80.0f;
      if (f.heading == -kPiF / 2) f.y = 80.0f; else if (f.heading == kPiF / 2) f.y = -80.0f;
      v = f;
    }
    oth[m] = v;
    crashed = crashed || hypotf(v.x - ego[0], v.y - ego[1]) < 2.5f;
  }
  const int t = a.t[b] + 1;
  const bool arrived = ego[0] <= a.arrive_x && fabsf(ego[1] - a.arrive_y) < 4.0f;
  const float speed_r = fminf(fmaxf((ego[3] - 7.0f) / 2.0f, 0.0f), 1.0f);   // lmap(speed, [7, 9], [0, 1]) clipped
  float reward = -5.0f * (crashed ? 1.0f : 0.0f) + speed_r;
  if (arrived) reward = 1.0f;
  const bool over = t >= a.duration_steps;
  const bool done = crashed || arrived || over;
  a.reward[b] = reward;
  a.done[b] = done;
  a.crashed[b] = crashed;
  a.arrived[b] = arrived;
  a.truncated[b] = over && !(crashed || arrived);
  a.speed[b] = ego[3];
  write_obs(a.terminal_obs + (size_t)b * V * 8, V, ego, oth, M);
  // ---- in-place reset of finished environments (draw calls 5-10; drawn for every env, used by the finished ones)
  if (done) {
    const uint64_t c5 = ctr + 4 * kCtrStep;
    const float r0 = u01(c5 + 1 * kCtrStep, b, 0), r1 = u01(c5 + 1 * kCtrStep, b, 1);
    ego[0] = 2.0f + 0.1f * gauss(c5 + 2 * kCtrStep, b, 0, 1);
    ego[1] = 50.0f + r0;                                   // envs/intersection_env__.py:303-315: spawns at the south entry
    ego[2] = -kPiF / 2;
    ego[3] = 8.0f + 2.0f * r1;
    for (int m = 0; m < M; ++m) {
      Veh v = spawn(c5 + 2 * kCtrStep, b, m, M);
      if (hypotf(v.x - ego[0], v.y - ego[1]) < 8.0f && v.heading == -kPiF / 2) v.y -= 20.0f;
      oth[m] = v;
    }
  }
  a.t[b] = done ? 0 : t;
  a.crashed_state[b] = done ? false : crashed;
  for (int i = 0; i < 4; ++i) a.ego[4 * b + i] = ego[i];
  for (int m = 0; m < M; ++m) {
    float* q = a.others + ((size_t)b * M + m) * 4;
    q[0] = oth[m].x; q[1] = oth[m].y; q[2] = oth[m].speed; q[3] = oth[m].heading;
  }
  write_obs(a.obs + (size_t)b * V * 8, V, ego, oth, M);
}

}  // namespace

extern "C" MPC_API int mpc_env_step(const MpcEnvStep* args, void* stream) {
  if (!args || args->B < 0 || args->n_others < 0 || args->n_others > kMaxOthers) return MPC_ERR_BAD_ARG;
  if (args->B == 0) return MPC_OK;
  if (!args->ego || !args->others || !args->t || !args->counter || !args->action || !args->obs || !args->terminal_obs ||
      !args->reward || !args->done || !args->crashed || !args->arrived || !args->truncated || !args->speed || !args->crashed_state)
    return MPC_ERR_BAD_ARG;
  k_env_step<<<(args->B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*args);
  return cudaGetLastError() == cudaSuccess ? MPC_OK : MPC_ERR_CUDA;
}
