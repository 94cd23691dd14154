/*
 * RoboCup physics substeps with the body state in registers.
 *
 * The substep loop of RoboCupEnvironment.step (reference paths under DynEnv/):
 *   processAction / Robot.step,turn,kick,turnHead   RoboCupEnvironment.py:527-581, Robot.py:103-138
 *   tick                                            :862-1007
 *   isBallOutOfField / ballFreeKickProcess          :622-732, :600-619
 *   pymunk Space.step + friction_robot / friction_ball   :482, cutils.py:93-140
 *
 * Lane b < 20 owns foot b (robot b / 2), lane 20 the ball, lanes 21..24 stand for the posts in the broad phase.  A
 * lane keeps its body -- position and angle as two-float numbers, velocity, bias velocity -- in registers for the whole
 * launch; both foot lanes of a robot keep identical copies of the robot's scalars (flags, timers, head, joint impulses,
 * reward sums) and update them with identical instructions, and the ball lane keeps the free-kick scalars.  The common
 * substep -- nobody falls, nobody is penalised, no pair is close or cached -- then runs on 21 lanes without touching
 * the record: partner feet talk through shuffles, only the group bounding circles of the broad phase and the cached
 * shape geometry go through shared memory.
 *
 * Everything rare is left to the record-based code of rc_sim.cuh, entered through a flush of the registers and left
 * through a reload: the ordered robot phase (falls, penalties, penalty re-entry, changes of the defenders lists;
 * detected BEFORE anything is modified, exactly like robot_phase_needs_order) and the contact half of Space.step
 * (space_step_rest, whenever the broad phase reports a candidate pair or a cached pair exists).  The arithmetic of the
 * fast path is the record-based code's, operation for operation: both produce identical bits.
 */
#pragma once
#include "rc_sim.cuh"
#include "rc_contact.cuh"

namespace dynenv {
namespace rc {

#ifdef RC_DBG_COUNTERS
__device__ unsigned long long g_dbg[16];
#define RC_DBG(k, c) do { if (E.lane == 0 && (c)) atomicAdd(&g_dbg[(SLOW ? 8 : 0) + (k)], 1ull); } while (0)
#define RC_DBGN(k, n) do { if (E.lane == 0) atomicAdd(&g_dbg[k], (unsigned long long)(n)); } while (0)
#else
#define RC_DBG(k, c)
#define RC_DBGN(k, n)
#endif

struct Body { /* lane-private: this lane's rigid body (lanes 21..24: px, py = the post's position) */
  float px, pxl, py, pyl, an, anl, vx, vy, w;
};
/* robot scalars, identical on both foot lanes; the ball lane keeps the free-kick state in the same registers:
   fl = ballOwned, mt = gracePeriod, hm = ballFreeCntr, prevx / prevy = the ball's prevPos.  What changes rarely stays
   in the record and is updated there: fall / penalty timers, reward sums, pending forces and bias velocities. */
struct Rob {
  int fl;
  float mt, hm, head, prevx, prevy, pivjx, pivjy, rotj;
};
struct Uni { /* warp-uniform */
  int closest0, closest1, elapsed;
};

__device__ __forceinline__ void regs_load(const Env E, Body& B, Rob& R, Uni& U) {
  const int lane = E.lane;
  if (lane <= BALL) {
    B.px = E.f[RC_S_PX + lane]; B.pxl = E.f[RC_S_PXL + lane]; B.py = E.f[RC_S_PY + lane]; B.pyl = E.f[RC_S_PYL + lane];
    B.an = E.f[RC_S_ANG + lane]; B.anl = E.f[RC_S_ANGL + lane];
    B.vx = E.f[RC_S_VX + lane]; B.vy = E.f[RC_S_VY + lane]; B.w = E.f[RC_S_W + lane];
  } else if (lane < NSHAPE) {
    const V2 pp = post_pos(lane - POST0);
    B.px = pp.x; B.py = pp.y;
  }
  if (lane < 20) {
    const int r = lane >> 1;
    R.fl = E.i[RC_S_RFLAGS + r]; R.mt = E.f[RC_S_MOVET + r]; R.hm = E.f[RC_S_HEADMOV + r]; R.head = E.f[RC_S_HEAD + r];
    R.prevx = E.f[RC_S_PREVX + r]; R.prevy = E.f[RC_S_PREVY + r];
    R.pivjx = E.f[RC_S_PIVJX + r]; R.pivjy = E.f[RC_S_PIVJY + r]; R.rotj = E.f[RC_S_ROTJ + r];
  } else if (lane == BALL) {
    R.fl = E.i[RC_S_BALLOWNED]; R.mt = E.f[RC_S_GRACE]; R.hm = E.f[RC_S_BALLFREE];
    R.prevx = E.f[RC_S_BPREVX]; R.prevy = E.f[RC_S_BPREVY];
  }
  U.closest0 = E.i[RC_S_CLOSEST]; U.closest1 = E.i[RC_S_CLOSEST + 1]; U.elapsed = E.i[RC_S_ELAPSED];
}

__device__ __forceinline__ void regs_store(const Env E, const Body& B, const Rob& R, const Uni& U) {
  const int lane = E.lane;
  if (lane <= BALL) {
    E.f[RC_S_PX + lane] = B.px; E.f[RC_S_PXL + lane] = B.pxl; E.f[RC_S_PY + lane] = B.py; E.f[RC_S_PYL + lane] = B.pyl;
    E.f[RC_S_ANG + lane] = B.an; E.f[RC_S_ANGL + lane] = B.anl;
    E.f[RC_S_VX + lane] = B.vx; E.f[RC_S_VY + lane] = B.vy; E.f[RC_S_W + lane] = B.w;
  }
  if (lane < 20 && !(lane & 1)) {
    const int r = lane >> 1;
    E.i[RC_S_RFLAGS + r] = R.fl; E.f[RC_S_MOVET + r] = R.mt; E.f[RC_S_HEADMOV + r] = R.hm; E.f[RC_S_HEAD + r] = R.head;
    E.f[RC_S_PREVX + r] = R.prevx; E.f[RC_S_PREVY + r] = R.prevy;
    E.f[RC_S_PIVJX + r] = R.pivjx; E.f[RC_S_PIVJY + r] = R.pivjy; E.f[RC_S_ROTJ + r] = R.rotj;
  } else if (lane == BALL) {
    E.i[RC_S_BALLOWNED] = R.fl; E.f[RC_S_GRACE] = R.mt; E.f[RC_S_BALLFREE] = R.hm;
    E.f[RC_S_BPREVX] = R.prevx; E.f[RC_S_BPREVY] = R.prevy;
  }
  if (lane == 0) { E.i[RC_S_CLOSEST] = U.closest0; E.i[RC_S_CLOSEST + 1] = U.closest1; E.i[RC_S_ELAPSED] = U.elapsed; }
  __syncwarp();
}

/* 10 substeps [sub_b, sub_e) of one environment; the record is in shared memory on entry and on exit.  Returns -1, or
   2 * sub + phase when substep `sub` needs the record-based code: phase 0 = from its start (ordered robot phase),
   phase 1 = from the narrow phase of its Space.step (a pair touches or is cached; positions are integrated).  The caller
   queues the environment for rc_phys_slow_kernel, which runs the rest of the launch's substeps (SLOW = true): keeping
   the rare code out of this kernel keeps its instruction footprint at the size of the common substep. */
/* SLOW = false: rc_phys_kernel, as above.  SLOW = true: rc_phys_slow_kernel -- the same register-resident substeps for
   the handed-over environments, which run the rare code in place: the ordered robot phase of rc_sim.cuh and the contact
   half of rc_contact.cuh, each between a flush and a reload of the registers; phase0 = 1 resumes inside Space.step of
   substep sub_b.  Returns -1. */
/* RC_MW (rc_slow.cu: several handed-over environments per block, one per warp): every warp of the block walks the
   substeps [sub_lo, sub_e) of the launch and meets the others at two block barriers per substep -- in front of the robot
   phase and in front of the velocity / contact half -- so that the warps of a block execute the same code at the same
   time and share its instruction fetches (the one-warp-per-block kernel issued at 28 % with 15 stalled-for-instruction
   warps per issue, whatever its occupancy).  A warp idles through the substeps in front of its environment's hand-over
   point (sub_b; phase0 = 1: it enters at the contact half of substep sub_b) and, without an environment (valid =
   false), through all of them. */
template <bool SLOW>
__device__ __forceinline__ int phys_substeps2(const Env E, const float* actions_env, int sub_b, int sub_e, bool first, int phase0 = 0,
                                              int sub_lo = 0, bool valid = true) {
  if (first) clear_step_scratch(E);
  const uint32_t pairs4 = broadphase_pairs(E);
#ifdef RC_MW
  static_assert(SLOW, "the multi-warp build holds only the handed-over environments' kernel");
#else
  (void)sub_lo; (void)valid;
  if (SLOW && phase0) {
    unsigned own, g0, g1, g2, g3;
    __syncwarp();
    const int prevStepped = E.i[RC_S_STEPPED];
    record_broadphase(E, pairs4, own, g0, g1, g2, g3);
    contact_half(E, own, g0, g1, g2, g3, pairs4, prevStepped);
    if (E.lane == 0) E.i[RC_S_ELAPSED] += 1;
    __syncwarp();
    sub_b++;
  }
#endif
  const int lane = E.lane;
  const int P = E.c->P, nPl = E.c->n_players;
  const bool isFoot = lane < 2 * P;       /* an existing robot's foot */
  const bool isBody = isFoot || lane == BALL;
  const int r = lane >> 1, side = lane & 1;
  const int teamI = r < nPl ? 1 : -1;
  const float dt = DT;
  __syncwarp();
  Body B = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  Rob R = {0, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  Uni U;
  regs_load(E, B, R, U);
  /* pending forces (fall) and bias velocities (contact position correction) live in the record: only the record-based
     code produces them; the next integration / velocity update reads them there */
  bool forcesPending = true, biasPending = true; /* (once per launch: the record-based code runs in another kernel) */
  const float penX = SIDE + PEN_LENGTH + LINEW / 2;

#pragma unroll 1
#ifdef RC_MW
  for (int sub = sub_lo; sub < sub_e; sub++) {
    __syncthreads();
    const bool act = valid && sub >= sub_b;            /* warp-uniform */
    const bool resume = act && phase0 && sub == sub_b; /* handed over in mid-substep: positions are integrated */
    const bool front = act && !resume;
#else
  for (int sub = sub_b; sub < sub_e; sub++) {
    const bool act = true, resume = false, front = true;
    (void)act;
#endif
    /* ------------------------------------------------------------------ robots: processAction (sub 0) + tick */
    if (front) {
      float4 act = make_float4(0, 0, 0, 0);
      Rand6 ru;
      ru.w[0] = ru.w[1] = ru.w[2] = ru.w[3] = 0;
      if (sub == 0 && isFoot) {
        act = *reinterpret_cast<const float4*>(actions_env + 4 * r);
        if (E.c->can_fall) ru = rng_block(E, (uint32_t)U.elapsed, RC_SITE_ACTION, r, 0, 0);
        if (bad_action(E, act)) E.i[X_DIAG] = E.i[X_DIAG] | RC_IF_BAD_ACTION; /* (the only writers now: same value) */
      }
      /* partner foot (hi words) and the robot's position */
      const float opx = __shfl_xor_sync(FULLMASK, B.px, 1), opy = __shfl_xor_sync(FULLMASK, B.py, 1);
      float posx = (B.px + opx) * 0.5f, posy = (B.py + opy) * 0.5f;
      const float ballx = __shfl_sync(FULLMASK, B.px, BALL), bally = __shfl_sync(FULLMASK, B.py, BALL);
      /* does this robot's phase touch anything but its own state?  (robot_phase_needs_order, on registers) */
      bool need = false;
      if (isFoot) {
        const int fl = R.fl;
        if (sub == 0 && E.c->can_fall && !(fl & (RC_RF_PENALIZED | RC_RF_KICKING | RC_RF_FALLEN)))
          need = (act.x > 0 && ru.u(0) > 0.999f) || (act.y > 0 && ru.u(1) > 0.999f) || (act.z > 0 && ru.u(2) > 0.99f);
        if ((fl & RC_RF_FALLEN) && E.f[RC_S_FALLT + r] - 10.f < 0) need = true;
        if (fl & RC_RF_JOINT_ADD_PENDING) need = true;
        if (posy < 0 || posx < 0 || posy > H_ || posx > W_) need = true; /* leaves the field: penalty */
        if (fl & RC_RF_PENALIZED) {
          if (E.f[RC_S_PENT + r] - 10.f <= 0) need = true;
        } else {
          const int teamIdx = teamI > 0 ? 0 : 1;
          const int nd = E.i[RC_S_NDEF + teamIdx];
          const bool inDef = (nd > 0 && E.i[RC_S_DEFENDERS + 2 * teamIdx] == r) || (nd > 1 && E.i[RC_S_DEFENDERS + 2 * teamIdx + 1] == r);
          {
            const float robX = teamIdx ? W_ - posx : posx;
            const bool inBox = robX < penX && posy > (H_ / 2 - PEN_WIDTH) && posy < (H_ / 2 + PEN_WIDTH);
            if (inBox != inDef) need = true;
          }
          if (fl & RC_RF_KICKING) { /* position with the kicking foot put back (the only own move of tick) */
            const bool kf = ((fl & RC_RF_FOOT) != 0) == (side == 1);
            const float ox = kf ? opx : B.px, oy = kf ? opy : B.py; /* the other foot */
            const float qx = (E.f[RC_S_INITX + r] + ox) * 0.5f, qy = (E.f[RC_S_INITY + r] + oy) * 0.5f;
            const float robX = teamIdx ? W_ - qx : qx;
            const bool inBox = robX < penX && qy > (H_ / 2 - PEN_WIDTH) && qy < (H_ / 2 + PEN_WIDTH);
            if (inBox != inDef) need = true;
            if (qy < 0 || qx < 0 || qy > H_ || qx > W_) need = true;
          }
        }
      }
      RC_DBG(0, true);
      if (__any_sync(FULLMASK, need)) {
        RC_DBG(1, true);
        regs_store(E, B, R, U);
        /* The register-resident kernel hands over -- except at substep 0, where the fall draws of processAction make 8 % of
           the environments need the ordered phase: those run it here, once, and stay (handed over they would walk their
           other nine substeps through the hand-over kernel's barriers: a third of its first launch of a step). */
        if (!SLOW && sub != 0) return 2 * sub;
        robots_phase(E, sub, actions_env);
        regs_load(E, B, R, U);
        forcesPending = true; biasPending = true;
      } else {
       bool restored = false;
       if (isFoot) {
        int fl = R.fl;
        if (sub == 0) { /* processAction without falls (RoboCupEnvironment.py:527-581, Robot.py:103-138) */
          const float move = act.x, turn = act.y, kick = act.z;
          float head = act.w;
          if (E.c->deterministic_turn) head = -3.f * teamI;
          if (!E.c->continuous) head -= 3.f;
          const bool canMove = !(fl & (RC_RF_PENALIZED | RC_RF_KICKING | RC_RF_FALLEN));
          if (move > 0 && canMove) {
            const int dir = (int)move - 1;
            R.mt = 500.f;
            if (side == 0) {
              V2 v = dir == 0 ? v2(0, 2 * ROB_VEL) : dir == 1 ? v2(0, -2 * ROB_VEL) : dir == 2 ? v2(2.5f * ROB_VEL, 0) : v2(-2 * ROB_VEL, 0);
              v = rot_ni(v, B.an);
              B.vx = v.x; B.vy = v.y;
            }
          }
          if (turn > 0 && canMove) {
            R.mt = 500.f;
            if (side == 0) B.w += ((int)turn - 1) ? ROB_ANGVEL : -ROB_ANGVEL;
          }
          if (head != 0) { /* Robot.turnHead */
            R.hm = head * PI_F / 720.f;
            R.mt = 500.f;
          }
          if (kick > 0 && move == 0 && turn == 0 && canMove) {
            const int foot = (int)kick - 1;
            if ((foot ? 1 : 0) == side) {
              E.f[RC_S_INITX + r] = B.px; E.f[RC_S_INITXL + r] = B.pxl;
              E.f[RC_S_INITY + r] = B.py; E.f[RC_S_INITYL + r] = B.pyl;
            }
            fl = (fl & ~RC_RF_FOOT) | RC_RF_KICKING | (foot ? RC_RF_FOOT : 0);
            R.mt = 1000.f;
          }
        }
        /* tick (RoboCupEnvironment.py:862-1007) */
        const float time = 10.f;
        float mt = R.mt;
        if (mt > 0) {
          mt -= time;
          if (R.hm != 0) R.head = fmaxf(-HEAD_MAX, fminf(HEAD_MAX, R.head + R.hm));
          if (fl & RC_RF_KICKING) {
            const bool kf = ((fl & RC_RF_FOOT) != 0) == (side == 1);
            if (mt + time > 500 && mt <= 500) {
              if (!(fl & RC_RF_JOINT_REMOVED)) fl |= RC_RF_JOINT_REMOVED | RC_RF_JOINT_ABSENT;
              if (kf) {
                V2 v = rot_ni(v2(ROB_VEL * 3.f, 0), B.an);
                B.vx = v.x; B.vy = v.y;
              }
            }
            if (mt + time > 400 && mt <= 400) {
              if (kf) {
                V2 v = rot_ni(v2(ROB_VEL * 2.5f, 0), B.an);
                B.vx = -v.x; B.vy = -v.y;
              }
            } else if (mt <= 300) {
              if (kf) {
                B.vx = 0.f; B.vy = 0.f;
                B.px = E.f[RC_S_INITX + r]; B.pxl = E.f[RC_S_INITXL + r];
                B.py = E.f[RC_S_INITY + r]; B.pyl = E.f[RC_S_INITYL + r];
              }
              restored = true;
              fl &= ~RC_RF_KICKING;
              if (fl & RC_RF_JOINT_REMOVED) fl &= ~(RC_RF_JOINT_REMOVED | RC_RF_JOINT_ABSENT);
            }
          }
          if (mt <= 0) {
            mt = 0;
            R.hm = 0.f;
            B.vx = 0.f; B.vy = 0.f; B.w = 0.f;
          }
          R.mt = mt;
        }
        R.fl = fl;
        if (!side) {
          if (fl & RC_RF_FALLEN) E.f[RC_S_FALLT + r] -= time; /* the end of the fall is an ordered event */
          if (fl & RC_RF_PENALIZED) E.f[RC_S_PENT + r] -= time; /* ... and so is the end of the penalty */
        }
       }
       /* the robot's position once more if a kicking foot was put back */
       if (__any_sync(FULLMASK, restored)) {
         const float qx = __shfl_xor_sync(FULLMASK, B.px, 1), qy = __shfl_xor_sync(FULLMASK, B.py, 1);
         posx = (B.px + qx) * 0.5f; posy = (B.py + qy) * 0.5f;
       }
       if (isFoot) {
        if (posx != R.prevx || posy != R.prevy) {
          if ((r == U.closest0 || r == U.closest1) && !(R.fl & RC_RF_PENALIZED)) {
            const float diff = len(v2(posx - ballx, posy - bally)) - len(v2(R.prevx - ballx, R.prevy - bally));
            if (!side) {
              E.f[X_REW + r] -= diff * 0.05f;
              E.f[X_POSREW + r] += fmaxf(0.f, -diff * 0.05f);
            }
          }
          R.prevx = posx; R.prevy = posy;
        }
       }
      }
      __syncwarp();
    }
    /* ------------------------------------------------------------------ ball rules (isBallOutOfField) */
    if (front) {
      float cr0 = 0, cr1 = 0, oldx = 0, oldy = 0;
      if (lane == BALL) {
        int finished = 0;
        const V2 pos = v2(B.px, B.py);
        oldx = pos.x; oldy = pos.y;
        const float outMin = SIDE - BALL_RADIUS, outMaxX = W_ - SIDE + BALL_RADIUS, outMaxY = H_ - SIDE + BALL_RADIUS;
        int team = 0;
        if (pos.y < outMin || pos.x < outMin || pos.y > outMaxY || pos.x > outMaxX) {
          float x = W_ / 2, y = H_ / 2;
          team = E.i[RC_S_NLASTK] ? (E.i[RC_S_LASTK] < nPl ? 1 : -1) : 1;
          if (pos.y < outMin || pos.y > outMaxY) {
            x = team < 0 ? pos.x + 50 : pos.x - 50;
            y = pos.y < outMin ? outMin + BALL_RADIUS : outMaxY - BALL_RADIUS;
          } else {
            if (pos.y < H_ / 2 + GOAL_W && pos.y > H_ / 2 - GOAL_W) {
              finished = 1;
              if (pos.x < outMin) { cr0 += -25; cr1 += 25; E.i[RC_S_GOALS + 1] += 1; }
              else { cr0 += 25; cr1 += -25; E.i[RC_S_GOALS] += 1; }
            } else {
              if (pos.x < outMin) {
                if (team < 0) x = SIDE + PEN_LENGTH;
                else { x = SIDE; y = pos.y < H_ / 2 ? SIDE : H_ - SIDE; }
              } else {
                if (team > 0) x = W_ - (SIDE + PEN_LENGTH);
                else { x = W_ - SIDE; y = pos.y < H_ / 2 ? SIDE : H_ - SIDE; }
              }
            }
          }
          B.px = x; B.py = y; B.pxl = 0.f; B.pyl = 0.f;
          B.vx = 0; B.vy = 0; B.w = 0;
        }
        /* ballFreeKickProcess(-team) (:600-619): fl = ballOwned, mt = gracePeriod, hm = ballFreeCntr */
        if (team == 0) {
          float g = R.mt;
          if (g > 0) {
            g -= 10.f;
            if (g < 0) { g = 0; R.hm = 9999.f; }
            R.mt = g;
          } else if (R.hm > 0) {
            float b = R.hm - 10.f;
            if (b < 0) { b = 0; R.fl = 0; }
            R.hm = b;
          }
        } else {
          R.fl = -team;
          R.mt = 14999.f;
          R.hm = 0.f;
        }
        const float bx = B.px;
        if (!finished) {
          const float d = (bx - R.prevx) / 20.f;
          cr0 += d;
          cr1 -= d;
        }
        R.prevx = bx;
        R.prevy = B.py;
        if (cr0 != 0.f || cr1 != 0.f) {
          E.f[X_TEAMREW] += cr0 * 0.1f;
          E.f[X_TEAMREW + 1] += cr1 * 0.1f;
        }
      }
      cr0 = __shfl_sync(FULLMASK, cr0, BALL);
      cr1 = __shfl_sync(FULLMASK, cr1, BALL);
      const float bx = __shfl_sync(FULLMASK, B.px, BALL), by = __shfl_sync(FULLMASK, B.py, BALL);
      const float opx = __shfl_xor_sync(FULLMASK, B.px, 1), opy = __shfl_xor_sync(FULLMASK, B.py, 1);
      const float rpx = (B.px + opx) * 0.5f, rpy = (B.py + opy) * 0.5f;
      if (cr0 != 0.f || cr1 != 0.f) { /* rewards of the ball's movement (zero currReward adds zeros) */
        oldx = __shfl_sync(FULLMASK, oldx, BALL);
        oldy = __shfl_sync(FULLMASK, oldy, BALL);
        if (isFoot && !side) {
          const int nlk = E.i[RC_S_NLASTK];
          float disc = 1.f, rsum = E.f[X_REW + r], psum = E.f[X_POSREW + r];
          bool inLast = false;
          for (int k = 0; k < nlk; k++) {
            if (E.i[RC_S_LASTK + k] == r) {
              const float rew = (r < nPl ? cr0 : cr1) * disc;
              rsum += rew;
              psum += fmaxf(0.f, rew);
              inLast = true;
            }
            disc *= 0.5f;
          }
          /* robots near the (old) ball position or closest to it share negative rewards */
          const bool cond1 = (r == U.closest0 || r == U.closest1);
          const bool cond2 = len(v2(rpx - oldx, rpy - oldy)) < 150.f;
          if ((cond1 || cond2) && !inLast) rsum += fminf(0.f, (r < nPl ? cr0 : cr1) * 0.5f);
          E.f[X_REW + r] = rsum; E.f[X_POSREW + r] = psum;
        }
      }
      /* closest robot per team: argmin of the squared distance, first minimum wins */
      {
        const float d2 = isFoot ? lensq(v2(bx - rpx, by - rpy)) : INFINITY;
        const unsigned bits = __float_as_uint(d2);
        const bool even = isFoot && side == 0;
        const unsigned m0 = __reduce_min_sync(FULLMASK, (even && r < nPl) ? bits : 0xffffffffu);
        const unsigned m1 = __reduce_min_sync(FULLMASK, (even && r >= nPl) ? bits : 0xffffffffu);
        const unsigned w0 = __ballot_sync(FULLMASK, even && r < nPl && bits == m0);
        const unsigned w1 = __ballot_sync(FULLMASK, even && r >= nPl && bits == m1);
        U.closest0 = w0 ? (__ffs(w0) - 1) >> 1 : 0;
        U.closest1 = w1 ? (__ffs(w1) - 1) >> 1 : nPl;
      }
    }
    /* ------------------------------------------------------------------ Space.step */
    {
      const int prevStepped = E.i[RC_S_STEPPED];
      const int nArb = E.i[RC_S_NARB];
      float bcx = 0, bcy = 0, dRLx = 0, dRLy = 0, dRLa = 0;
      unsigned own = 0, g0 = 0, g1 = 0, g2 = 0, g3 = 0;
      if (front) {
      /* integrate positions (cpBodyUpdatePosition), refresh the cached shape geometry */
      float tax = 0, tay = 0, tbx = 0, tby = 0;
      if (isBody) {
        float vbx = 0.f, vby = 0.f, wb = 0.f;
        if (biasPending) {
          vbx = E.f[RC_S_VBX + lane]; vby = E.f[RC_S_VBY + lane]; wb = E.f[RC_S_WB + lane];
          E.f[RC_S_VBX + lane] = 0.f; E.f[RC_S_VBY + lane] = 0.f; E.f[RC_S_WB + lane] = 0.f;
        }
        two_add(B.px, B.pxl, __fmul_rn(B.vx + vbx, dt));
        two_add(B.py, B.pyl, __fmul_rn(B.vy + vby, dt));
        two_add(B.an, B.anl, __fmul_rn(B.w + wb, dt));
        V2 ta, tb;
        shape_world(lane, B.px, B.py, B.an, &ta, &tb);
        E.f[RC_S_CAX + lane] = ta.x; E.f[RC_S_CAY + lane] = ta.y;
        E.f[RC_S_CBX + lane] = tb.x; E.f[RC_S_CBY + lane] = tb.y;
        bcx = (ta.x + tb.x) * 0.5f; bcy = (ta.y + tb.y) * 0.5f;
        tax = ta.x; tay = ta.y; tbx = tb.x; tby = tb.y;
      } else if (lane < NSHAPE) {
        bcx = B.px; bcy = B.py;
      }
      biasPending = false;
      /* partner foot, both words of every coordinate */
      const float opx = __shfl_xor_sync(FULLMASK, B.px, 1), opxl = __shfl_xor_sync(FULLMASK, B.pxl, 1);
      const float opy = __shfl_xor_sync(FULLMASK, B.py, 1), opyl = __shfl_xor_sync(FULLMASK, B.pyl, 1);
      const float oan = __shfl_xor_sync(FULLMASK, B.an, 1), oanl = __shfl_xor_sync(FULLMASK, B.anl, 1);
      /* right foot minus left foot, the same bits on both lanes (pos_diff(R, L)) */
      const float Rx = side ? B.px : opx, Rxl = side ? B.pxl : opxl, Lx = side ? opx : B.px, Lxl = side ? opxl : B.pxl;
      const float Ry = side ? B.py : opy, Ryl = side ? B.pyl : opyl, Ly = side ? opy : B.py, Lyl = side ? opyl : B.pyl;
      const float Ra = side ? B.an : oan, Ral = side ? B.anl : oanl, La = side ? oan : B.an, Lal = side ? oanl : B.anl;
      dRLx = __fadd_rn(__fsub_rn(Rx, Lx), __fsub_rn(Rxl, Lxl));
      dRLy = __fadd_rn(__fsub_rn(Ry, Ly), __fsub_rn(Ryl, Lyl));
      dRLa = __fadd_rn(__fsub_rn(Ra, La), __fsub_rn(Ral, Lal));
      /* broad phase on groups, lane = pair, four rounds.  A robot's group circle: every point of a foot (a capsule
         (-10, +-10)-(10, +-10) of radius 7.5 around its body position) lies within sqrt(200) + 7.5 of the foot's body
         position, whatever its angle, and the body positions within half their distance of their midpoint.  (Penalised
         robots parked on neighbouring spots, 52.5 apart, do not overlap with this radius; with the 27.5 of the first
         version they did, on every substep.) */
      const float ROB_BOUND = 21.65f;
      bool ownFeetApart = false;
      if (lane < 20) {
        if (!side) {
          const float ddx = B.px - opx, ddy = B.py - opy;
          E.f[X_GX + r] = (B.px + opx) * 0.5f;
          E.f[X_GY + r] = (B.py + opy) * 0.5f;
          E.f[X_GR + r] = ROB_BOUND + 0.5f * sqrtf(ddx * ddx + ddy * ddy);
        }
      } else if (lane < NSHAPE) {
        E.f[X_GX + lane - 10] = bcx; E.f[X_GY + lane - 10] = bcy; E.f[X_GR + lane - 10] = 10.f;
      }
      /* a robot's own feet can only touch if an end point of the partner's segment comes within the two radii of this
         foot's supporting line (normally it runs 20 away, also while a kicking foot slides along its axis) */
      {
        const float oax = __shfl_xor_sync(FULLMASK, tax, 1), oay = __shfl_xor_sync(FULLMASK, tay, 1);
        const float obx = __shfl_xor_sync(FULLMASK, tbx, 1), oby = __shfl_xor_sync(FULLMASK, tby, 1);
        if (isFoot && side) {
          const float ex = tbx - tax, ey = tby - tay; /* |e| = 20 */
          const float c1 = ex * (oay - tay) - ey * (oax - tax), c2 = ex * (oby - tay) - ey * (obx - tax);
          /* the right foot's segment has the left one on its positive side, 20 away: cross = +400 */
          ownFeetApart = !(c1 > 20.f * (2 * ROB_RAD + 0.1f) && c2 > 20.f * (2 * ROB_RAD + 0.1f));
        }
      }
      __syncwarp();
      own = __ballot_sync(FULLMASK, ownFeetApart);
      {
        unsigned gh[4];
#pragma unroll
        for (int rnd = 0; rnd < 4; rnd++) {
          const int a = (pairs4 >> (8 * rnd + 4)) & 15, b = (pairs4 >> (8 * rnd)) & 15;
          const float dx = E.f[X_GX + a] - E.f[X_GX + b], dy = E.f[X_GY + a] - E.f[X_GY + b];
          const float rr = E.f[X_GR + a] + E.f[X_GR + b] + 0.05f;
          gh[rnd] = __ballot_sync(FULLMASK, dx * dx + dy * dy < rr * rr);
        }
        g0 = gh[0]; g1 = gh[1]; g2 = gh[2]; g3 = gh[3];
      }
      } else if (resume) {
        __syncwarp();
        record_broadphase(E, pairs4, own, g0, g1, g2, g3);
      }
#ifdef RC_MW
      __syncthreads();
      if (!act) continue;
#endif
      RC_DBG(2, own); RC_DBG(3, g0 | g1 | g2 | g3); RC_DBG(4, nArb);
      if (SLOW) {
        if (resume | own | g0 | g1 | g2 | g3 | (unsigned)nArb) { /* a pair is close or cached: the contact half runs on the record */
          RC_DBG(6, true);
          if (!resume) {
            if (lane < NSHAPE) { E.f[X_BCX + lane] = bcx; E.f[X_BCY + lane] = bcy; }
            regs_store(E, B, R, U);
          }
          contact_half(E, own, g0, g1, g2, g3, pairs4, prevStepped);
          regs_load(E, B, R, U);
          forcesPending = true; biasPending = true;
          U.elapsed += 1;
          __syncwarp();
          continue;
        }
      } else {
        /* groups that came close: exact test of their shape pairs on the cached geometry */
        int kept = 0;
        if (own | g0 | g1 | g2 | g3) {
          if (lane < NSHAPE) { E.f[X_BCX + lane] = bcx; E.f[X_BCY + lane] = bcy; }
          __syncwarp();
          kept = filter_candidates(E, own, g0, g1, g2, g3);
        }
        RC_DBG(5, kept);
        if (kept | nArb) { /* a pair touches or is cached: the contact half of the substep runs on the record */
          RC_DBG(6, true);
          regs_store(E, B, R, U);
          return 2 * sub + 1;
        }
      }
      {
        if (!prevStepped && lane == 0) E.i[RC_S_STEPPED] = 1;
        const float biasCoef = 6.1259606640e-02f; /* 1 - collisionBias^dt, collisionBias = 0.9^60 */
        /* velocity update through the friction callbacks (cutils.py:93-140) */
        if (isBody) {
          const float minv = body_minv(lane), iinv = body_iinv(lane), m = body_m(lane);
          const float friction = lane == BALL ? 2.8e-2f : 1e-3f, rotFriction = lane == BALL ? 1e-3f : 1e-2f;
          const float spin = lane == BALL ? 5e-2f : 0.f;
          float x = B.vx, y = B.vy, theta = B.w;
          if (forcesPending) {
            x = x + (E.f[RC_S_FX + lane] * minv) * dt;
            y = y + (E.f[RC_S_FY + lane] * minv) * dt;
            theta = theta + E.f[RC_S_TQ + lane] * iinv * dt;
            E.f[RC_S_FX + lane] = 0.f; E.f[RC_S_FY + lane] = 0.f; E.f[RC_S_TQ + lane] = 0.f;
          }
          const float factor = friction * m, rotFactor = rotFriction * m;
          const float length = 1.f / (fabsf(x) + fabsf(y) + 1e-5f);
          float a0 = x * factor * length, a1 = y * factor * length;
          a0 += a1 * spin * theta;
          a1 -= a0 * spin * theta;
          x = fabsf(x) < factor ? 0.f : x - a0;
          y = fabsf(y) < factor ? 0.f : y - a1;
          theta = fabsf(theta) < rotFactor ? 0.f : theta - (theta > 0 ? rotFactor : -rotFactor);
          B.vx = x; B.vy = y; B.w = theta;
        }
        forcesPending = false;
        /* joints (lane = foot, both feet of a robot compute the same impulses): pivot with both anchors at the centres
           of gravity, rotary limit 0..0; warm start, then the sweeps (two: without contacts every joint is an
           independent one-constraint system that the first sweep solves exactly) */
        const float ovx = __shfl_xor_sync(FULLMASK, B.vx, 1), ovy = __shfl_xor_sync(FULLMASK, B.vy, 1), ow = __shfl_xor_sync(FULLMASK, B.w, 1);
        if (isFoot) {
          const float pivK = 1.f / (2 * FOOT_MINV), rotISum = 1.f / (2 * FOOT_IINV);
          const bool pivOn = !(R.fl & RC_RF_JOINT_ABSENT);
          float pivBx = 0, pivBy = 0;
          if (pivOn) {
            const float c = -2.2762779044e+00f; /* -(1 - errorBias^dt) / dt, joint.error_bias = 0.1 (Robot.py:59) */
            pivBx = dRLx * c;
            pivBy = dRLy * c;
          }
          const float pdist = -dRLa; /* min = max = 0 */
          const float rotBias = -biasCoef * pdist / dt;
          if (rotBias == 0) R.rotj = 0;
          const float dt_coef = prevStepped ? 1.f : 0.f;
          float lvx = side ? ovx : B.vx, lvy = side ? ovy : B.vy, lw = side ? ow : B.w;
          float rvx = side ? B.vx : ovx, rvy = side ? B.vy : ovy, rw = side ? B.w : ow;
          if (pivOn) {
            const float jx = R.pivjx * dt_coef, jy = R.pivjy * dt_coef;
            lvx -= jx * FOOT_MINV; lvy -= jy * FOOT_MINV;
            rvx += jx * FOOT_MINV; rvy += jy * FOOT_MINV;
          }
          {
            const float j = R.rotj * dt_coef;
            lw -= j * FOOT_IINV;
            rw += j * FOOT_IINV;
          }
#pragma unroll
          for (int it = 0; it < 2; it++) {
            if (pivOn) { /* cpPivotJoint applyImpulse, r1 = r2 = 0 */
              const float jx = pivK * (pivBx - (rvx - lvx)), jy = pivK * (pivBy - (rvy - lvy));
              R.pivjx += jx; R.pivjy += jy;
              lvx -= jx * FOOT_MINV; lvy -= jy * FOOT_MINV;
              rvx += jx * FOOT_MINV; rvy += jy * FOOT_MINV;
            }
            if (rotBias != 0) { /* cpRotaryLimitJoint applyImpulse */
              float j = -(rotBias + (rw - lw)) * rotISum;
              const float jOld = R.rotj;
              R.rotj = rotBias < 0 ? fmaxf(jOld + j, 0.f) : fminf(jOld + j, 0.f);
              j = R.rotj - jOld;
              lw -= j * FOOT_IINV;
              rw += j * FOOT_IINV;
            }
          }
          B.vx = side ? rvx : lvx; B.vy = side ? rvy : lvy; B.w = side ? rw : lw;
        }
      }
    }
    U.elapsed += 1;
    __syncwarp();
  }
  regs_store(E, B, R, U);
  return -1;
}

}  // namespace rc
}  // namespace dynenv
