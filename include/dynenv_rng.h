/*
 * RNG stream layout shared by the CUDA kernels and the CPU oracle.
 *
 * The reference draws from Python's global `random` / `np.random` sequentially
 * (environment_base.py:190-193).  The batched simulator uses the counter-based
 * Philox4x32-10 generator (Salmon et al., SC'11) instead, so that every draw is addressed
 * by WHERE it is consumed and results do not depend on execution order or GPU count:
 *
 *   key     = (seed low 32, seed high 32)
 *   counter = (global env id, episode, tick, site<<24 | a<<16 | b<<8 | block)
 *   uniform k of a site = 21-bit slot (k % 6) of block (k / 6): the block's four words
 *   w0..w3 are read as one 128-bit little-endian integer, slot s = bits [21 s, 21 s + 21),
 *   u = slot * 2^-21 in [0,1).  Six uniforms per Philox evaluation is what every noise
 *   routine of the observation path needs (cutils.py:352-468), so each sighting costs ONE
 *   Philox block; 21 bits are exact in float32, so the float32 kernels and the float64
 *   oracle see identical uniforms.
 *
 * tick: substep index `elapsed` (before the increment) for draws made inside the substep
 * loop; `elapsed` at snapshot time for observation ticks; DYNENV_TICK_RESETOBS | t for the
 * nTimeSteps initial observations; DYNENV_TICK_RESET for scene generation.
 */
#ifndef DYNENV_RNG_H
#define DYNENV_RNG_H

#define DYNENV_TICK_RESET 0xFFFFFFFFu
#define DYNENV_TICK_RESETOBS 0x80000000u

/* RoboCup sites (a, b = the two 8-bit indices) */
#define RC_SITE_ACTION 1   /* a = robot; k: 0 move fall, 1 turn fall, 2 kick fall  (RoboCupEnvironment.py:557,565,577) */
#define RC_SITE_REFALL 2   /* a = robot; k: 0                                      (:932) */
#define RC_SITE_ROBCOL 3   /* a, b = shape ids of the pair; k: 0 robot1, 1 robot2  (:1064,1068) */
#define RC_SITE_POSTCOL 4  /* a, b = shape ids of the pair; k: 0                   (:1121) */
#define RC_SITE_RESET 5    /* a = 0; k: draws of _create_robot_spots/_create_agents/_create_ball in call order (:239-336) */
/* observation sites: a = observer, b = object index
   k for addNoise (cutils.py:417-468): 0,1 position, 2 false negative, 3 misclassification, 4 size, 5 angle
   k for addNoiseLine (cutils.py:352-378): 0..3 end-point noise, 4 false negative */
#define RC_SITE_OBS_BALL 16
#define RC_SITE_OBS_ROBOT 17
#define RC_SITE_OBS_POST 18
#define RC_SITE_OBS_CROSS 19
#define RC_SITE_OBS_FCROSS 20
#define RC_SITE_OBS_CIRCLE 21
#define RC_SITE_OBS_LINE 22
#define RC_SITE_OBS_MISCROSS 23 /* k: 0,1 the two randint(-1,1) tags of a ball misclassified as cross (:1260) */
#define RC_SITE_OBS_FP 24       /* b = loop index 0..9; k: 0 test, 1 class, 2 distance, 3 angle, 4.. class draws (:1279-1305) */
#define RC_SITE_OBS_FPBALL 25   /* b = row in the robot list; k: 0 test, 1 drop test, 2,3 offset, 4 size (:1308-1316) */


/* Driving sites */
#define DRV_SITE_RESET 32      /* k: draws of _create_agents / _create_pedestrians / _create_obstacles in call order (DrivingEnvironment.py:76-98,527-584) */
#define DRV_SITE_PEDMOVE 33    /* a = pedestrian; k: 0 duration, 1 speed, 2 crossing chance, 3 crossing speed (:470-500) */
/* observation sites: a = observer, b = entity index.  addNoiseRect (cutils.py:479-542): k 0,1 position, 2 false negative,
   3 misclassification, 4 angle; addNoiseLane (:380-413): k 0 distance, 1 angle, 2 false negative */
#define DRV_SITE_OBS_SELF 40
#define DRV_SITE_OBS_CAR 41
#define DRV_SITE_OBS_PED 42
#define DRV_SITE_OBS_OBST 43
#define DRV_SITE_OBS_LANE 44
#define DRV_SITE_OBS_FP 45     /* b = loop index; k: 0 test, 1 class, 2 distance, 3 direction, 4 object angle, 5.. class draws (:828-874) */
#define DRV_SITE_OBS_FPPED 46  /* b = row of the car list; k: 0 test, 1, 2 offset (:877-882) */

#endif
