/*
 * dynenv_b200 — C ABI of the B200-native batched simulator for DynEnv's rollout hot path.
 *
 * Plain C, no torch types: device pointers + sizes only.  Every entry point returns 0 on
 * success or a negative DYNENV_E_* code; dynenv_last_error() gives the message.  Calls are
 * asynchronous on the given CUDA stream (no hidden device synchronisation) unless stated.
 * One handle owns the persistent simulation state of nEnvs environments on one GPU.
 *
 * What each entry point replaces in the reference (szemenyeim/DynEnv, paths under DynEnv/):
 *   dynenv_create       RoboCupEnvironment.__init__  RoboCupEnvironment.py:23-64
 *                       DrivingEnvironment.__init__  DrivingEnvironment.py:20-56
 *                       make_dyn_env                 __init__.py:6-25 (num_envs -> nEnvs)
 *   dynenv_reset        EnvironmentBase.reset        environment_base.py:205-224
 *                       + auto-reset of SubprocVecEnv._worker  utils/subproc_vec_env.py:17-23
 *   dynenv_step         RoboCupEnvironment.step      RoboCupEnvironment.py:446-524
 *                       DrivingEnvironment.step      DrivingEnvironment.py:248-322
 *                       (the pymunk Space.step inside it: environment_base.py:126-128,
 *                        RoboCupEnvironment.py:482, DrivingEnvironment.py:278)
 *   dynenv_set_seed     EnvironmentBase.set_random_seed  environment_base.py:190-193
 *   dynenv_get_state /  no reference equivalent (the reference pokes pymunk bodies directly);
 *   dynenv_set_state    needed for state injection in parity tests
 */
#ifndef DYNENV_B200_H
#define DYNENV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DYNENV_ABI_VERSION 2

/* ---- error codes */
#define DYNENV_OK 0
#define DYNENV_E_INVALID (-1)   /* bad argument / configuration */
#define DYNENV_E_CUDA (-2)      /* CUDA runtime error */
#define DYNENV_E_NODEVICE (-3)  /* no CUDA device: there is no CPU fallback */
#define DYNENV_E_UNSUPPORTED (-4)

/* ---- enums (values follow DynEnv/cutils.py:10-64,152-165) */
#define DYNENV_KIND_ROBOCUP 0
#define DYNENV_KIND_DRIVING 1
#define DYNENV_OBS_FULL 0
#define DYNENV_OBS_PARTIAL 1
#define DYNENV_OBS_IMAGE 2
#define DYNENV_NOISE_RANDOM 0
#define DYNENV_NOISE_REALISTIC 1

/* class-level toggles of RoboCupEnvironment (RoboCupEnvironment.py:18-21) */
#define DYNENV_FLAG_RANDOM_INIT 1u
#define DYNENV_FLAG_DETERMINISTIC_TURN 2u
#define DYNENV_FLAG_CAN_FALL 4u
#define DYNENV_FLAG_USE_OBS_REWARDS 8u
#define DYNENV_FLAG_AUTO_RESET 16u /* vec-env semantics: done envs are re-initialised in step */
/* RoboCup PARTIAL observations as a counts-prefixed stream: a record's rows BEHIND its counts (RC_O_COUNTS) are left as
   the caller's buffer had them instead of being zero-filled -- a consumer that honours the counts (dynenv_arrange_inputs,
   the reference's own list format) reads the same rows, and the step writes ~25 KB of observations per env-step instead
   of 83 KB (70 % of a padded record is padding).  Off: every record is fully written, rows behind the counts are zero. */
#define DYNENV_FLAG_SPARSE_OBS 32u

typedef struct DynEnvConfig {
  int32_t abi_version;      /* DYNENV_ABI_VERSION */
  int32_t kind;             /* DYNENV_KIND_* */
  int32_t n_envs;           /* environments owned by this handle */
  int32_t n_players;        /* per team (RoboCup, clamped to 5) / cars (Driving, clamped to 10) */
  int32_t obs_type;         /* DYNENV_OBS_* */
  int32_t noise_type;       /* DYNENV_NOISE_* */
  float noise_magnitude;    /* 0..5 */
  int32_t continuous;       /* RoboCup: allowHeadTurn, Driving: continuousActions */
  uint32_t flags;           /* DYNENV_FLAG_* */
  uint64_t seed;            /* Philox key */
  int64_t env_id_offset;    /* global id of env 0 (multi-GPU sharding: results do not depend on
                               how many GPUs the env range is split over) */
  int32_t device;           /* CUDA device ordinal */
} DynEnvConfig;

typedef struct DynEnvHandle_s* DynEnvHandle;

/* ---- RoboCup fixed capacities (rows per observer and tick, after false positives) */
#define RC_MAXP 10        /* robots per env (2 teams x 5) */
#define RC_NBODY 21       /* 20 feet + ball; body 2r = left foot of robot r, 2r+1 = right, 20 = ball */
#define RC_NTICK 5        /* observation ticks per step (substeps 9,19,29,39,49) */
#define RC_SUBSTEPS 50
#define RC_CAP_BALL 8
#define RC_CAP_ROBOT 12
#define RC_CAP_POST 6
#define RC_CAP_CROSS 6
#define RC_CAP_FCROSS 18
#define RC_CAP_LINE 11
#define RC_F_BALL 5
#define RC_F_ROBOT 7
#define RC_F_POST 6
#define RC_F_CROSS 6
#define RC_F_FCROSS 8
#define RC_F_LINE 5
#define RC_NTYPES 6       /* ball, robot, post, cross, fcross, line */
#define RC_MAXARB 32      /* cached contact pairs per env */

/*
 * RoboCup persistent state: one record of RC_STATE_WORDS 32-bit words per environment,
 * records contiguous ([nEnvs][RC_STATE_WORDS]) so that the warp that owns an environment
 * loads/stores it with fully coalesced 128-byte transactions.  F = float32, I = int32.
 * (Appendix D of SURVEY.md lists where each field lives in the reference.)
 */
enum {
  /* bodies, arrays of RC_NBODY */
  RC_S_PX = 0,                       /* F position x          */
  RC_S_PY = RC_S_PX + RC_NBODY,      /* F position y          */
  RC_S_ANG = RC_S_PY + RC_NBODY,     /* F angle               */
  RC_S_VX = RC_S_ANG + RC_NBODY,     /* F velocity            */
  RC_S_VY = RC_S_VX + RC_NBODY,
  RC_S_W = RC_S_VY + RC_NBODY,       /* F angular velocity    */
  RC_S_FX = RC_S_W + RC_NBODY,       /* F pending force       */
  RC_S_FY = RC_S_FX + RC_NBODY,
  RC_S_TQ = RC_S_FY + RC_NBODY,      /* F pending torque      */
  RC_S_VBX = RC_S_TQ + RC_NBODY,     /* F bias velocity (position correction, consumed next substep) */
  RC_S_VBY = RC_S_VBX + RC_NBODY,
  RC_S_WB = RC_S_VBY + RC_NBODY,
  /* world geometry of the shapes as of the last substep's collision phase (feet: segment end
     points a, b; ball: centre in a) — what point queries made before the next substep see */
  RC_S_CAX = RC_S_WB + RC_NBODY,
  RC_S_CAY = RC_S_CAX + RC_NBODY,
  RC_S_CBX = RC_S_CAY + RC_NBODY,
  RC_S_CBY = RC_S_CBX + RC_NBODY,
  /* robots, arrays of RC_MAXP (Robot.py:58-88) */
  RC_S_HEAD = RC_S_CBY + RC_NBODY,   /* F headAngle   */
  RC_S_HEADMOV = RC_S_HEAD + RC_MAXP,/* F headMoving  */
  RC_S_MOVET = RC_S_HEADMOV + RC_MAXP,/* F moveTime   */
  RC_S_INITX = RC_S_MOVET + RC_MAXP, /* F initPos (kick) */
  RC_S_INITY = RC_S_INITX + RC_MAXP,
  RC_S_FALLT = RC_S_INITY + RC_MAXP, /* F fallTime    */
  RC_S_PENT = RC_S_FALLT + RC_MAXP,  /* F penalTime   */
  RC_S_PREVX = RC_S_PENT + RC_MAXP,  /* F prevPos     */
  RC_S_PREVY = RC_S_PREVX + RC_MAXP,
  RC_S_PIVJX = RC_S_PREVY + RC_MAXP, /* F pivot joint accumulated impulse */
  RC_S_PIVJY = RC_S_PIVJX + RC_MAXP,
  RC_S_ROTJ = RC_S_PIVJY + RC_MAXP,  /* F rotary-limit joint accumulated impulse */
  RC_S_RFLAGS = RC_S_ROTJ + RC_MAXP, /* I RC_RF_* bits */
  RC_S_FALLCNT = RC_S_RFLAGS + RC_MAXP, /* I fallCntr */
  RC_S_TOUCHCNT = RC_S_FALLCNT + RC_MAXP, /* I touchCntr */
  RC_S_EPR = RC_S_TOUCHCNT + RC_MAXP,  /* F episodeRewards    */
  RC_S_EPPR = RC_S_EPR + RC_MAXP,      /* F episodePosRewards */
  RC_S_EPOR = RC_S_EPPR + RC_MAXP,     /* F episodeObsRewards */
  /* ball (Ball.py:20-24) */
  RC_S_BPREVX = RC_S_EPOR + RC_MAXP, /* F */
  RC_S_BPREVY = RC_S_BPREVX + 1,     /* F */
  RC_S_LASTK = RC_S_BPREVY + 1,      /* I lastKicked[4] */
  RC_S_NLASTK = RC_S_LASTK + 4,      /* I */
  /* env scalars (RoboCupEnvironment.py:42-59) */
  RC_S_ELAPSED = RC_S_NLASTK + 1,    /* I */
  RC_S_EPISODE = RC_S_ELAPSED + 1,   /* I episode counter (RNG stream id) */
  RC_S_BALLOWNED = RC_S_EPISODE + 1, /* I -1/0/1 */
  RC_S_BALLFREE = RC_S_BALLOWNED + 1,/* F ballFreeCntr */
  RC_S_GRACE = RC_S_BALLFREE + 1,    /* F gracePeriod  */
  RC_S_GOALS = RC_S_GRACE + 1,       /* I [2] */
  RC_S_CLOSEST = RC_S_GOALS + 2,     /* I closestID[2] */
  RC_S_DEFENDERS = RC_S_CLOSEST + 2, /* I defenders[2][2] */
  RC_S_NDEF = RC_S_DEFENDERS + 4,    /* I [2] */
  RC_S_PENALTIMES = RC_S_NDEF + 2,   /* I penalTimes[2] */
  RC_S_STEPPED = RC_S_PENALTIMES + 2,/* I 0 until the first substep after (re)initialisation (warm-start gate) */
  /* cached contact pairs (Chipmunk arbiters): RC_MAXARB entries */
  RC_S_NARB = RC_S_STEPPED + 1,      /* I */
  RC_S_ARBKEY = RC_S_NARB + 1,       /* I a | b<<8 | state<<16 | age<<24   (shape ids: body id, posts 21..24) */
  RC_S_ARBCON = RC_S_ARBKEY + RC_MAXARB, /* I ncontacts | hash0<<8 | hash1<<16 */
  RC_S_ARBJN0 = RC_S_ARBCON + RC_MAXARB, /* F */
  RC_S_ARBJT0 = RC_S_ARBJN0 + RC_MAXARB,
  RC_S_ARBJN1 = RC_S_ARBJT0 + RC_MAXARB,
  RC_S_ARBJT1 = RC_S_ARBJN1 + RC_MAXARB,
  /* low-order parts of positions and angles: a coordinate is the unevaluated sum hi + lo of two
     float32 words (RC_S_PX + RC_S_PXL, ...), which keeps the small differences the joint and
     contact solvers work on (foot-to-foot offsets, penetration depths) exact to ~1e-13 although
     all arithmetic is float32.  Consumers that do not need that read the hi word alone. */
  RC_S_PXL = RC_S_ARBJT1 + RC_MAXARB,   /* F [RC_NBODY] */
  RC_S_PYL = RC_S_PXL + RC_NBODY,
  RC_S_ANGL = RC_S_PYL + RC_NBODY,
  RC_S_INITXL = RC_S_ANGL + RC_NBODY,   /* F [RC_MAXP] */
  RC_S_INITYL = RC_S_INITXL + RC_MAXP,
  /* accumulators of the step in flight: a step is a sequence of kernel launches (substep blocks,
     observation ticks, finish), which hand these over through the record.  Zeroed at the first
     substep of a step; meaningless between steps (the oracle leaves them 0). */
  RC_S_STEP_REW = RC_S_INITYL + RC_MAXP,        /* F [RC_MAXP] robotRewards so far            */
  RC_S_STEP_POSREW = RC_S_STEP_REW + RC_MAXP,   /* F [RC_MAXP] robotPosRewards                */
  RC_S_STEP_TEAMREW = RC_S_STEP_POSREW + RC_MAXP, /* F [2] teamRewards                        */
  RC_S_STEP_DIAG = RC_S_STEP_TEAMREW + 2,       /* I RC_IF_* flags                            */
  RC_S_STEP_LSUM = RC_S_STEP_DIAG + 1,          /* I [RC_MAXP] sum over ticks of numLandMarks */
  RC_S_STEP_BSUM = RC_S_STEP_LSUM + RC_MAXP,    /* I [RC_MAXP] sum of ballsSeen               */
  RC_S_STEP_RSEEN = RC_S_STEP_BSUM + RC_MAXP,   /* I [RC_MAXP][RC_MAXP-1] sum of robotsSeen   */
  RC_S_STEP_END = RC_S_STEP_RSEEN + RC_MAXP * (RC_MAXP - 1),
  RC_STATE_WORDS_RAW = RC_S_STEP_END,
  RC_STATE_WORDS = (RC_STATE_WORDS_RAW + 31) / 32 * 32 /* records start 128-byte aligned */
};

/* robot flag bits (RC_S_RFLAGS) */
#define RC_RF_KICKING 1
#define RC_RF_FOOT 2          /* 1 = right foot kicks */
#define RC_RF_FALLEN 4
#define RC_RF_PENALIZED 8
#define RC_RF_TOUCHING 16
#define RC_RF_MIGHTPUSH 32
#define RC_RF_JOINT_REMOVED 64 /* robot.jointRemoved (logical flag) */
#define RC_RF_JOINT_ABSENT 128 /* pivot joint currently not in the space (differs from the flag
                                  while a re-add issued inside a collision callback is deferred) */
#define RC_RF_JOINT_ADD_PENDING 256

/* cached-pair states (Chipmunk cpArbiterState) */
#define RC_ARB_FIRST 0
#define RC_ARB_NORMAL 1
#define RC_ARB_IGNORE 2
#define RC_ARB_CACHED 3

/*
 * Step/reset outputs: caller-owned device buffers (NULL = not wanted), written as packed
 * 32-bit-word records so that the warp owning an environment stores them with fully
 * coalesced 128-byte transactions.  E = nEnvs, T = RC_NTICK, P = 2*n_players.
 * Observation rows beyond the count are zero.  Row orders follow the reference's list
 * orders (RoboCupEnvironment.py:1192-1561).  F = float32, I = int32.
 *
 * PARTIAL observation record, one per (env, tick, observer): obs[E][T][P][RC_OBS_WORDS]
 * (what getAgentVision returns, RoboCupEnvironment.py:1537-1561)
 */
enum {
  RC_O_BALL = 0,                                       /* F [RC_CAP_BALL][RC_F_BALL]     */
  RC_O_ROBOT = RC_O_BALL + RC_CAP_BALL * RC_F_BALL,    /* F [RC_CAP_ROBOT][RC_F_ROBOT]   */
  RC_O_POST = RC_O_ROBOT + RC_CAP_ROBOT * RC_F_ROBOT,  /* F [RC_CAP_POST][RC_F_POST]     */
  RC_O_CROSS = RC_O_POST + RC_CAP_POST * RC_F_POST,    /* F [RC_CAP_CROSS][RC_F_CROSS]   */
  RC_O_FCROSS = RC_O_CROSS + RC_CAP_CROSS * RC_F_CROSS,/* F [RC_CAP_FCROSS][RC_F_FCROSS] */
  RC_O_LINE = RC_O_FCROSS + RC_CAP_FCROSS * RC_F_FCROSS,/* F [RC_CAP_LINE][RC_F_LINE]    */
  RC_O_COUNTS = RC_O_LINE + RC_CAP_LINE * RC_F_LINE,   /* I [RC_NTYPES] rows per type    */
  RC_O_NLANDMARKS = RC_O_COUNTS + RC_NTYPES,           /* I numLandMarks (before false positives) */
  RC_O_BALLSEEN = RC_O_NLANDMARKS + 1,                 /* I ballsSeen                    */
  RC_O_ROBSEEN = RC_O_BALLSEEN + 1,                    /* I robotsSeen[RC_MAXP-1] (first P-1 used) */
  RC_OBS_WORDS_RAW = RC_O_ROBSEEN + RC_MAXP - 1,
  RC_OBS_WORDS = (RC_OBS_WORDS_RAW + 31) / 32 * 32
};

/* FULL observation record (getFullState(robot), RoboCupEnvironment.py:1162-1188), one per
 * (env, tick, observer) in full[E][T][P][RC_FULL_WORDS] when obs_type == FULL, and one per
 * (env, observer) inside the info record (info['Recon States']) */
enum {
  RC_U_BALL = 0,                 /* F [4]            */
  RC_U_SELF = 4,                 /* F [8]            */
  RC_U_OTHERS = 12,              /* F [RC_MAXP-1][6] (first P-1 rows used) */
  RC_FULL_WORDS = 12 + (RC_MAXP - 1) * 6 + 2 /* 68 */
};

/* per-environment result record: info[E][RC_INFO_WORDS] */
enum {
  RC_I_REWARDS = 0,                         /* F [RC_MAXP] robotRewards                       */
  RC_I_DONE = RC_I_REWARDS + RC_MAXP,       /* I finished                                     */
  RC_I_EPR = RC_I_DONE + 1,                 /* F [RC_MAXP] episode_r   (valid where done)     */
  RC_I_EPPR = RC_I_EPR + RC_MAXP,           /* F [RC_MAXP] episode_p_r                        */
  RC_I_EPOR = RC_I_EPPR + RC_MAXP,          /* F [RC_MAXP] episode_o_r                        */
  RC_I_EPG = RC_I_EPOR + RC_MAXP,           /* I [2] episode_g                                */
  RC_I_FLAGS = RC_I_EPG + 2,                /* I RC_IF_* diagnostics of this step             */
  RC_I_STATE_ROBOTS = RC_I_FLAGS + 1,       /* F [RC_MAXP][6] info['Full State'][0]           */
  RC_I_STATE_BALL = RC_I_STATE_ROBOTS + RC_MAXP * 6, /* F [3] info['Full State'][1]           */
  RC_I_RECON = RC_I_STATE_BALL + 3,         /* F [RC_MAXP][RC_FULL_WORDS] info['Recon States'] */
  RC_INFO_WORDS_RAW = RC_I_RECON + RC_MAXP * RC_FULL_WORDS,
  RC_INFO_WORDS = (RC_INFO_WORDS_RAW + 31) / 32 * 32
};
#define RC_IF_ARB_OVERFLOW 1    /* more than RC_MAXARB cached contact pairs: a pair was dropped   */
#define RC_IF_ACTIVE_OVERFLOW 2 /* more than RC_MAXACTIVE touching pairs in one substep           */
#define RC_IF_BAD_ACTION 4      /* an action of this step was out of range or not finite: the reference raises
                                   (RoboCupEnvironment.py:543-550, DrivingEnvironment.py:364-367); the batched device path cannot,
                                   so it flags the step (the move / turn / kick / steering code then did something, but
                                   nothing the reference defines) */
#define RC_MAXACTIVE 16

/*
 * IMAGE observations (obs_type IMAGE; the image branch of getAgentVision, RoboCupEnvironment.py:1318-1467,1533-1534).
 * The observation kernel emits, per (env, tick, observer), the detection lists the reference draws -- positions and
 * sizes in the observer's head frame -- as one raw record raw[E][T][P][RC_RAW_WORDS]; dynenv_render_images turns raw
 * records into the reference's 8 x 480 x 640 planes (top camera 0..3, bottom camera 4..7; channel 0 balls, 1 robots,
 * 2 goal posts, 3 lines / centre circle / crosses), one bit per pixel: image[record][8][480][RC_IMG_WORDS_PER_ROW],
 * bit x % 32 of word x / 32 = pixel x.  An environment-step of 5 ticks x 10 observers is 15.4 MB of planes, so the
 * caller renders the records it needs (any sub-range) into a buffer of its own size.
 */
enum {
  RC_R_COUNTS = 0,                                   /* I [RC_NTYPES] rows per list (as RC_O_COUNTS)       */
  RC_R_CIRCLE_SEEN = RC_R_COUNTS + RC_NTYPES,        /* I centre circle sighted                             */
  RC_R_CIRCLE = RC_R_CIRCLE_SEEN + 1,                /* F [3] x, y, radius                                  */
  RC_R_BALL = RC_R_CIRCLE + 3,                       /* F [RC_CAP_BALL][3] x, y, size                       */
  RC_R_ROBOT = RC_R_BALL + RC_CAP_BALL * 3,          /* F [RC_CAP_ROBOT][3]                                 */
  RC_R_POST = RC_R_ROBOT + RC_CAP_ROBOT * 3,         /* F [RC_CAP_POST][3]                                  */
  RC_R_CROSS = RC_R_POST + RC_CAP_POST * 3,          /* F [RC_CAP_CROSS][3]                                 */
  RC_R_FCROSS = RC_R_CROSS + RC_CAP_CROSS * 3,       /* F [RC_CAP_FCROSS][3]                                */
  RC_R_LINE = RC_R_FCROSS + RC_CAP_FCROSS * 3,       /* F [RC_CAP_LINE][4] x1, y1, x2, y2                   */
  RC_RAW_WORDS_RAW = RC_R_LINE + RC_CAP_LINE * 4,
  RC_RAW_WORDS = (RC_RAW_WORDS_RAW + 31) / 32 * 32
};
#define RC_IMG_PLANES 8
#define RC_IMG_H 480
#define RC_IMG_W 640
#define RC_IMG_WORDS_PER_ROW 20
#define RC_IMG_WORDS (RC_IMG_PLANES * RC_IMG_H * RC_IMG_WORDS_PER_ROW) /* per (env, tick, observer): 307 200 bytes */

typedef struct DynEnvRoboCupOut {
  float* obs;   /* [E][T][P][RC_OBS_WORDS]  PARTIAL / IMAGE observations (obs_type PARTIAL or IMAGE), else NULL */
  float* full;  /* [E][T][P][RC_FULL_WORDS] FULL observations (obs_type FULL), else NULL       */
  float* info;  /* [E][RC_INFO_WORDS]       rewards, done, episode stats, full/recon state     */
  float* raw;   /* [E][T][P][RC_RAW_WORDS]  raw detection records (obs_type IMAGE), else NULL  */
} DynEnvRoboCupOut;

/* ======================================================================= Driving
 * DrivingEnvironment (DrivingEnvironment.py:20-56): N cars (boxes), up to 20 pedestrians (circles), up to
 * 20 obstacles and 4 buildings (static boxes), two roads.  Shape ids follow the reference's insertion
 * order (_setup_scene :58-63): buildings 0..3, cars 4..13, pedestrians 14..33, obstacles 34..53. */
#define DRV_MAXCARS 10
#define DRV_MAXPED 20
#define DRV_MAXOBS 20
#define DRV_NBUILD 4
#define DRV_SUBSTEPS 10
#define DRV_MAXARB 32
#define DRV_MAXACTIVE 16
#define DRV_SHAPE_CAR0 4
#define DRV_SHAPE_PED0 14
#define DRV_SHAPE_OBS0 34
#define DRV_NSHAPE 54
#define DRV_NLANEROWS 6     /* getCarLaneDistances rows: 4 (road 0) + 2 (road 1) */
/* observation capacities per observer (after false positives) */
#define DRV_CAP_CAR 14
#define DRV_CAP_OBS 24
#define DRV_CAP_PED 28
#define DRV_CAP_LANE 10
#define DRV_F_CAR 7
#define DRV_F_OBS 6
#define DRV_F_PED 2
#define DRV_F_LANE 4
#define DRV_F_SELF 9

enum {
  /* cars, arrays of DRV_MAXCARS (Car.py:15-52) */
  DRV_S_CPX = 0,                              /* F position (hi word, see DRV_S_CPXL)  */
  DRV_S_CPY = DRV_S_CPX + DRV_MAXCARS,
  DRV_S_CANG = DRV_S_CPY + DRV_MAXCARS,       /* I angle: HIGH 32 bits of the float64 (low: DRV_S_CANGL) */
  DRV_S_CVX = DRV_S_CANG + DRV_MAXCARS,       /* F velocity                            */
  DRV_S_CVY = DRV_S_CVX + DRV_MAXCARS,
  DRV_S_CW = DRV_S_CVY + DRV_MAXCARS,         /* F angular velocity                    */
  DRV_S_CVBX = DRV_S_CW + DRV_MAXCARS,        /* F bias velocity                       */
  DRV_S_CVBY = DRV_S_CVBX + DRV_MAXCARS,
  DRV_S_CWB = DRV_S_CVBY + DRV_MAXCARS,
  DRV_S_CDIRX = DRV_S_CWB + DRV_MAXCARS,      /* F car.direction                       */
  DRV_S_CDIRY = DRV_S_CDIRX + DRV_MAXCARS,
  DRV_S_CPREVX = DRV_S_CDIRY + DRV_MAXCARS,   /* F prevPos                             */
  DRV_S_CPREVY = DRV_S_CPREVX + DRV_MAXCARS,
  DRV_S_CGOALX = DRV_S_CPREVY + DRV_MAXCARS,  /* F goal                                */
  DRV_S_CGOALY = DRV_S_CGOALX + DRV_MAXCARS,
  DRV_S_CTYPE = DRV_S_CGOALY + DRV_MAXCARS,   /* I type 0..3 (mass, size, power)       */
  DRV_S_CTEAM = DRV_S_CTYPE + DRV_MAXCARS,    /* I                                     */
  DRV_S_CFLAGS = DRV_S_CTEAM + DRV_MAXCARS,   /* I DRV_CF_*                            */
  DRV_S_CLANE = DRV_S_CFLAGS + DRV_MAXCARS,   /* I car.position (LanePosition)         */
  /* pedestrians, arrays of DRV_MAXPED (Pedestrian.py:21-37) */
  DRV_S_PPX = DRV_S_CLANE + DRV_MAXCARS,
  DRV_S_PPY = DRV_S_PPX + DRV_MAXPED,
  DRV_S_PVX = DRV_S_PPY + DRV_MAXPED,
  DRV_S_PVY = DRV_S_PVX + DRV_MAXPED,
  DRV_S_PVBX = DRV_S_PVY + DRV_MAXPED,
  DRV_S_PVBY = DRV_S_PVBX + DRV_MAXPED,
  DRV_S_PMOVING = DRV_S_PVBY + DRV_MAXPED,    /* F moving (ms left)                    */
  DRV_S_PSPEED = DRV_S_PMOVING + DRV_MAXPED,  /* I speed 3..6                          */
  DRV_S_PFLAGS = DRV_S_PSPEED + DRV_MAXPED,   /* I DRV_PF_*                            */
  /* obstacles, arrays of DRV_MAXOBS (half extents 10 x 10) */
  DRV_S_OPX = DRV_S_PFLAGS + DRV_MAXPED,
  DRV_S_OPY = DRV_S_OPX + DRV_MAXOBS,
  /* environment scalars */
  DRV_S_ELAPSED = DRV_S_OPY + DRV_MAXOBS,     /* I */
  DRV_S_EPISODE = DRV_S_ELAPSED + 1,          /* I */
  DRV_S_ALLFIN = DRV_S_EPISODE + 1,           /* I allFinished */
  DRV_S_NPED = DRV_S_ALLFIN + 1,              /* I */
  DRV_S_NOBS = DRV_S_NPED + 1,                /* I */
  DRV_S_STEPPED = DRV_S_NOBS + 1,             /* I */
  DRV_S_EPR = DRV_S_STEPPED + 1,              /* F [DRV_MAXCARS] episodeRewards    */
  DRV_S_EPPR = DRV_S_EPR + DRV_MAXCARS,       /* F [DRV_MAXCARS] episodePosRewards */
  /* cached contact pairs that involve a car (pedestrian-pedestrian and pedestrian-obstacle pairs are
     ignored by the reference's handlers and never stored) */
  DRV_S_NARB = DRV_S_EPPR + DRV_MAXCARS,
  DRV_S_ARBKEY = DRV_S_NARB + 1,              /* I a | b<<8 | state<<16 | age<<24 */
  DRV_S_ARBCON = DRV_S_ARBKEY + DRV_MAXARB,   /* I ncontacts | hash0<<8 | hash1<<16 */
  DRV_S_ARBJN0 = DRV_S_ARBCON + DRV_MAXARB,
  DRV_S_ARBJT0 = DRV_S_ARBJN0 + DRV_MAXARB,
  DRV_S_ARBJN1 = DRV_S_ARBJT0 + DRV_MAXARB,
  DRV_S_ARBJT1 = DRV_S_ARBJN1 + DRV_MAXARB,
  /* low-order words of the two-float coordinates */
  DRV_S_CPXL = DRV_S_ARBJT1 + DRV_MAXARB,
  DRV_S_CPYL = DRV_S_CPXL + DRV_MAXCARS,
  DRV_S_CANGL = DRV_S_CPYL + DRV_MAXCARS,     /* I low 32 bits of the float64 car angles */
  DRV_S_PPXL = DRV_S_CANGL + DRV_MAXCARS,
  DRV_S_PPYL = DRV_S_PPXL + DRV_MAXPED,
  /* accumulators of the step in flight */
  DRV_S_STEP_REW = DRV_S_PPYL + DRV_MAXPED,       /* F [DRV_MAXCARS] carRewards    */
  DRV_S_STEP_POSREW = DRV_S_STEP_REW + DRV_MAXCARS,
  DRV_S_STEP_TEAMREW = DRV_S_STEP_POSREW + DRV_MAXCARS,
  DRV_S_STEP_DIAG = DRV_S_STEP_TEAMREW + 1,
  DRV_S_STEP_END = DRV_S_STEP_DIAG + 1,
  DRV_STATE_WORDS_RAW = DRV_S_STEP_END,
  DRV_STATE_WORDS = (DRV_STATE_WORDS_RAW + 31) / 32 * 32
};
#define DRV_CF_FINISHED 1
#define DRV_CF_CRASHED 2
#define DRV_PF_DEAD 1
#define DRV_PF_CROSSING 2
#define DRV_PF_BEGINCROSSING 4
#define DRV_PF_SIDE 8
#define DRV_PF_ROAD 16     /* 0: the vertical road, 1: the horizontal one */

/* PARTIAL observation record per (env, observer): obs[E][N][DRV_OBS_WORDS]
 * ((cars, obstacles, pedestrians), (self, lanes), (1,1,1)) of getAgentVision (DrivingEnvironment.py:960-977);
 * FULL observations (get_full_obs :121-124) use the same record: cars = the other cars, obstacles rows hold
 * 4 floats, lanes rows 5 floats (DRV_F_LANE_FULL), no noise. */
#define DRV_F_LANE_FULL 5
enum {
  DRV_O_SELF = 0,                                           /* F [DRV_F_SELF]              */
  DRV_O_CAR = DRV_O_SELF + DRV_F_SELF + 3,                  /* F [DRV_CAP_CAR][DRV_F_CAR]  */
  DRV_O_OBS = DRV_O_CAR + DRV_CAP_CAR * DRV_F_CAR,          /* F [DRV_CAP_OBS][DRV_F_OBS]  */
  DRV_O_PED = DRV_O_OBS + DRV_CAP_OBS * DRV_F_OBS,          /* F [DRV_CAP_PED][DRV_F_PED]  */
  DRV_O_LANE = DRV_O_PED + DRV_CAP_PED * DRV_F_PED,         /* F [DRV_CAP_LANE][DRV_F_LANE_FULL] */
  DRV_O_COUNTS = DRV_O_LANE + DRV_CAP_LANE * DRV_F_LANE_FULL, /* I [4] cars, obstacles, pedestrians, lanes */
  DRV_OBS_WORDS_RAW = DRV_O_COUNTS + 4,
  DRV_OBS_WORDS = (DRV_OBS_WORDS_RAW + 31) / 32 * 32
};
/* per-environment result record: info[E][DRV_INFO_WORDS] */
enum {
  DRV_I_REWARDS = 0,                         /* F [DRV_MAXCARS] carRewards */
  DRV_I_DONE = DRV_I_REWARDS + DRV_MAXCARS,  /* I */
  DRV_I_EPR = DRV_I_DONE + 1,                /* F [DRV_MAXCARS] */
  DRV_I_EPPR = DRV_I_EPR + DRV_MAXCARS,
  DRV_I_EPG = DRV_I_EPPR + DRV_MAXCARS,      /* I [2] finished-and-not-crashed, crashed */
  DRV_I_FLAGS = DRV_I_EPG + 2,               /* I diagnostics (RC_IF_*) */
  DRV_I_STATE_CARS = DRV_I_FLAGS + 1,        /* F [DRV_MAXCARS][7] info['Full State'][0] */
  DRV_I_STATE_OBS = DRV_I_STATE_CARS + DRV_MAXCARS * 7, /* F [DRV_MAXOBS][4] */
  DRV_I_STATE_PED = DRV_I_STATE_OBS + DRV_MAXOBS * 4,   /* F [DRV_MAXPED][2] */
  DRV_I_NOBS = DRV_I_STATE_PED + DRV_MAXPED * 2,        /* I */
  DRV_I_NPED = DRV_I_NOBS + 1,                          /* I */
  DRV_I_GOALS = DRV_I_NPED + 1,                         /* F [DRV_MAXCARS][2] goals, normalised like the state rows
                                                           (the self row of info['Recon States']) */
  DRV_INFO_WORDS_RAW = DRV_I_GOALS + DRV_MAXCARS * 2,
  DRV_INFO_WORDS = (DRV_INFO_WORDS_RAW + 31) / 32 * 32
};

typedef struct DynEnvDrivingOut {
  float* obs;   /* [E][N][DRV_OBS_WORDS] */
  float* info;  /* [E][DRV_INFO_WORDS]   */
} DynEnvDrivingOut;

/* ---- entry points */
int dynenv_abi_version(void);
const char* dynenv_last_error(void);

int dynenv_create(const DynEnvConfig* cfg, DynEnvHandle* out);
int dynenv_destroy(DynEnvHandle h);

/* words per environment record of this handle's state (RC_STATE_WORDS / DRV_STATE_WORDS) */
int dynenv_state_words(DynEnvHandle h);

/* Re-initialise environments (all if env_mask_dev == NULL, else those with mask != 0) and
   write the initial observations (nTimeSteps ticks of the initial scene) into `out`.
   `out` points to a DynEnvRoboCupOut / DynEnvDrivingOut according to the handle's kind. */
int dynenv_reset(DynEnvHandle h, const uint8_t* env_mask_dev, void* out, void* stream);

/* One env step for every environment.  actions_dev: RoboCup float32 [E][P][4]
   (move, turn, kick, head), Driving float32 [E][N][2].  With DYNENV_FLAG_AUTO_RESET,
   `terminal_out` (same struct type, may be NULL) receives the observations of environments
   that finished this step while `out` receives their post-reset observations. */
/* (dynenv_step_u8: the same step for MultiDiscrete action spaces -- RoboCup without allowHeadTurn, Driving without
   continuousActions (RoboCupEnvironment.py:400-408, DrivingEnvironment.py:200-207) -- with the action indices as bytes,
   uint8 [E][P][4] / [E][N][2]: a quarter of the host -> device traffic of a step.  The bytes are widened to the handle's own
   float32 buffer by a kernel on `stream`, then the step runs as below; a value a float action could not have is flagged like
   one (RC_IF_BAD_ACTION).  The first call allocates that buffer: make it before capturing the step into a CUDA graph --
   under capture it returns DYNENV_E_INVALID instead.) */
int dynenv_step_u8(DynEnvHandle h, const uint8_t* actions_u8_dev, void* out, void* terminal_out, void* stream);
int dynenv_step(DynEnvHandle h, const float* actions_dev, void* out, void* terminal_out,
                void* stream);

int dynenv_set_seed(DynEnvHandle h, uint64_t seed);

/* IMAGE observations: draw n_records raw records (raw_dev: [n_records][RC_RAW_WORDS], e.g. any slice of
   DynEnvRoboCupOut.raw) into image_dev: [n_records][RC_IMG_WORDS] uint32 (layout above).  Replaces the drawing half
   of getAgentVision (RoboCupEnvironment.py:1318-1467) and cutils.projectPoints / estimateConic / getConicPoints
   (cutils.py:206-284).  `device`: CUDA device ordinal of the buffers. */
int dynenv_render_images(int device, const float* raw_dev, int64_t n_records, uint32_t* image_dev, void* stream);

/* Rollout buffers on the device (models/storage.py): what the learner keeps of every env.step() result, without the host
   round trip of models/train.py:269-284.  All buffers are the caller's.
   dynenv_rollout_insert -- storage.insert (models/storage.py:133-168) for the environment-side fields, from the result record
   `out` that dynenv_step just filled: rewards_dev[step][nEnvs * nAgents], agent_finished_dev[step][nEnvs * nAgents] (the last
   column of every agent's row of info['Full State'][0], models/train.py:276), dones_dev[step][nEnvs]; and
   storage.log_episode_rewards (:292-308) as running sums per environment in episode_acc_dev[nEnvs][DYNENV_EPISODE_ACC_WORDS]
   (may be NULL): [0] episodes finished, [1..3] sums of the agent-mean episode_r / episode_p_r / episode_o_r, [4..5] sums of
   episode_g[0..1], [6..7] min and max of the agent-mean episode_r, then a ring of the last DYNENV_EPISODE_RING episodes'
   per-agent episode_r, [ring slot][DYNENV_EPISODE_RING_AGENTS] (slot = episode number % DYNENV_EPISODE_RING; the
   reference's deque keeps the last 10 * num_envs entries, models/storage.py:46, for the median of print_reward_stats
   :310-318).  The three step-indexed pointers are the buffers' bases.
   dynenv_discount_rewards -- storage._discount_rewards (models/storage.py:170-205): out[i] = rewards[i] + discount * out[i+1],
   out[T] = final_value, over n_columns = nEnvs * nAgents columns.  dones_dev NULL reproduces the reference (its mask is built
   from zeros, :194); with dones_dev ([T][n_columns / agents_per_env]) the chain restarts at a step that ended an episode. */
#define DYNENV_EPISODE_RING 10
#define DYNENV_EPISODE_RING_AGENTS 10
#define DYNENV_EPISODE_STAT_WORDS 8
#define DYNENV_EPISODE_ACC_WORDS (DYNENV_EPISODE_STAT_WORDS + DYNENV_EPISODE_RING * DYNENV_EPISODE_RING_AGENTS)
int dynenv_rollout_insert(DynEnvHandle h, const void* out, int step, int rollout_size, float* rewards_dev,
                          uint8_t* agent_finished_dev, uint8_t* dones_dev, float* episode_acc_dev, void* stream);
int dynenv_discount_rewards(int device, const float* rewards_dev, const float* final_value_dev, const uint8_t* dones_dev,
                            int agents_per_env, float* out_dev, int rollout_size, int64_t n_columns, float discount,
                            void* stream);

/* The learner-side re-shuffle of the observation records on the device: InOutArranger (models/models.py:127-274).
   A layout describes where each object type's rows and count sit in one observer's record (rec[E][T][P][rec_words]).
   dynenv_arrange_inputs  -- rearrange_inputs (models.py:220-247): flat_out[k] = all rows of type k in (time step,
     environment x player, row) order ([rows][feat[k]]; capacity: T * E * P * cap[k] rows, or NULL to skip the type),
     counts_out[n_types][T][E*P], obj_counts_out[T][E*P], offsets_out[n_types][T][E*P] + block_sums_out[n_types][nb]
     (nb = ceil(T*E*P / 1024); row of (k, t, q) = offsets[k][t*Q+q] + block_sums[k][(t*Q+q) / 1024]), totals_out[n_types + 1]
     (rows per type; [n_types] = maxCount).  Nothing returns to the host.
   dynenv_arrange_outputs -- rearrange_outputs (models.py:249-274): emb[k] = the type's embeddings [rows][F] (NULL: the type has no
     block) -> padded_out[T][max_count][E*P][F] (per player its rows of type 0, then type 1, ..., zeros behind) and
     mask_out[T][E*P][max_count] (1 = padding, may be NULL).  max_count: the caller's static bound (>= totals[n_types]). */
#define DYNENV_ARRANGE_MAXTYPES 8
typedef struct DynEnvArrangeLayout {
  int32_t n_types, rec_words;
  int32_t off[DYNENV_ARRANGE_MAXTYPES];        /* first word of the type's rows in the record */
  int32_t cap[DYNENV_ARRANGE_MAXTYPES];        /* rows the record can hold */
  int32_t feat[DYNENV_ARRANGE_MAXTYPES];       /* floats per row handed on */
  int32_t stride[DYNENV_ARRANGE_MAXTYPES];     /* floats per row in the record (>= feat) */
  int32_t count_word[DYNENV_ARRANGE_MAXTYPES]; /* word of the record holding the row count (int32); -1: always one row */
} DynEnvArrangeLayout;
int dynenv_arrange_inputs(int device, const float* rec_dev, int64_t n_envs, int n_time, int n_players, const DynEnvArrangeLayout* layout,
                          float* const* flat_out, int32_t* counts_out, int32_t* obj_counts_out, int32_t* offsets_out,
                          int32_t* block_sums_out, int32_t* totals_out, void* stream);
int dynenv_arrange_outputs(int device, const float* const* emb, int n_features, int n_types, int64_t n_envs, int n_time, int n_players,
                           const int32_t* counts, const int32_t* offsets, const int32_t* block_sums, const int32_t* obj_counts,
                           int max_count, float* padded_out, uint8_t* mask_out, void* stream);

/* copy the whole state ([nEnvs][state_words] 32-bit words) to / from a device buffer */
int dynenv_get_state(DynEnvHandle h, uint32_t* state_dev, void* stream);
int dynenv_set_state(DynEnvHandle h, const uint32_t* state_dev, void* stream);

/* Measurement aid: with timing enabled, every launch of the following dynenv_step calls is bracketed by
   CUDA events on the caller's stream; dynenv_last_kernel_ms waits for the last step's launches and returns
   the summed duration (ms) and the launch count per kernel class (arrays of FOUR entries): [0] substep blocks
   (rc_phys_kernel / drv_phys_kernel), [1] observation ticks (rc_obsp_kernel / rc_obs_kernel / drv_obs_kernel),
   [2] finish (rc_finish_kernel), [3] substep blocks of the environments handed over to the record-based code
   (rc_phys_slow_kernel).  enable = 1: the step runs as always -- the handed-over environments on the side stream beside
   the others' observations -- and [3] is what the side stream adds beyond them; enable = 2: the side stream's work is
   issued on the caller's stream instead, one kernel after the other, and [3] is the hand-over kernel's own duration (the
   per-class times then add up like an ncu launch list; the step is slower than in production). */
int dynenv_debug_timing(DynEnvHandle h, int enable);
int dynenv_last_kernel_ms(DynEnvHandle h, float* ms4, int* launches4);

/* Measurement aid: FP32 (CUDA-core, non-tensor) FMA throughput of `device` in TFLOP/s, from a register-resident FMA
   kernel timed with CUDA events (8 independent chains per thread, 148 x 8 blocks of 256 threads, warmed up): the
   denominator SURVEY 8d asks for next to the step kernels' flop rate.  Synchronises the device. */
int dynenv_debug_fp32_peak(int device, float* tflops_out);

/* number of kernels launched by the last dynenv_step / dynenv_reset call on this handle */
int dynenv_last_launch_count(DynEnvHandle h);

/* Debugging aid for parity work: the following dynenv_step calls run only substeps [begin, end) of
   the step (actions are applied at substep 0, observations are emitted at substeps 9, 19, ...;
   rewards of a partial step cover only the substeps run).  Default 0, RC_SUBSTEPS. */
int dynenv_debug_substep_range(DynEnvHandle h, int begin, int end);

#ifdef __cplusplus
}
#endif
#endif /* DYNENV_B200_H */
