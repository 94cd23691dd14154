"""Drop-in API of the reference for the rollout path (host side, Python like the reference's):

  RoboCupEnvironment(nPlayers, render, observationType, noiseType, noiseMagnitude, obs_space_cast, allowHeadTurn)
      .step(actions) .reset() .set_random_seed(seed) .observation_space .action_space ...
      (DynEnv/RoboCupEnvironment.py:23-64,446-524; environment_base.py:190-224)
  DrivingEnvironment(nPlayers, render, observationType, noiseType, noiseMagnitude, obs_space_cast, continuousActions)
      (DynEnv/DrivingEnvironment.py:20-56,248-322)
  make_dyn_env(env, num_envs, num_players, render, observationType, noiseType, noiseMagnitude,
               use_continuous_actions) -> (vec_env, name)                     (DynEnv/__init__.py:6-25)
  BatchedVecEnv: the VecEnv calls the learner uses (step / step_async / step_wait / reset / get_attr /
      env_method / seed / close; DynEnv/utils/base_vec_env.py:57-233, subproc_vec_env.py:50-175) over ONE
      batched simulator instead of one process per environment.

Same names, argument meaning and error behaviour.  The simulation itself runs in the CUDA library; these
classes only validate, launch and convert formats."""
from dataclasses import dataclass
from typing import Any, List, Tuple

import numpy as np

from . import layout as L
from . import reference_format as RF
from . import spaces as S
from .enums import DynEnvType, NoiseType, ObservationType

_ACT_HI = (5, 3, 3, 7)


# ---- reconstruction descriptors (DynEnv/environment_base.py:18-38): what `env.recoDescriptor` /
# `vec_env.get_attr('recoDescriptor', 0)` hand to the reconstruction head of the learner (examples/main.py:42)
@dataclass
class StateSpaceDescriptor:
    numItemsPerGridCell: int
    space: Any


@dataclass
class PredictionDescriptor:
    numContinuous: int = None  # position not included
    numBinary: int = 0         # confidence not included
    contIdx: List[int] = None
    binaryIdx: List[int] = None
    posIdx: List[int] = (0, 1)
    categoricIdx: int = None


@dataclass
class RecoDescriptor:
    featureGridSize: Tuple[int, int]
    fullStateSpace: List[StateSpaceDescriptor]
    targetDefs: List[PredictionDescriptor]


def robocup_reco_descriptor():
    """RoboCupEnvironment._setup_reconstruction_info (RoboCupEnvironment.py:101-132): ball and robot items on a 1 x 1 grid"""
    xy, conf = S.Box(-2, +2, shape=(2,)), S.MultiBinary(1)
    ball = StateSpaceDescriptor(1, S.Dict({"position": xy, "team": S.Box(-1, 1, shape=(1,)), "confidence": conf}))
    robot = StateSpaceDescriptor(4, S.Dict({"position": xy, "orientation": S.Box(-1.0, +1.0, shape=(2,)),
                                            "team": S.Box(-1, 1, shape=(1,)), "active": S.MultiBinary(1), "confidence": conf}))
    return RecoDescriptor(featureGridSize=(1, 1), fullStateSpace=[ball, robot],
                          targetDefs=[PredictionDescriptor(numContinuous=1, contIdx=[2, ]),
                                      PredictionDescriptor(numContinuous=3, numBinary=1, contIdx=[2, 3, 4], binaryIdx=[5, ])])


def driving_reco_descriptor(mean=5.0):
    """DrivingEnvironment._setup_reconstruction_info (DrivingEnvironment.py:129-168): self, cars, obstacles, pedestrians on
    a 10 x 17 grid"""
    xy, ori, size, conf = S.Box(-mean * 2, +mean * 2, shape=(2,)), S.Box(-1.0, +1.0, shape=(2,)), S.Box(-10, 10, shape=(2,)), S.MultiBinary(1)
    full = {"position": xy, "orientation": ori, "size": size, "confidence": conf}
    return RecoDescriptor(featureGridSize=(10, 17),
                          fullStateSpace=[StateSpaceDescriptor(1, S.Dict(dict(full))), StateSpaceDescriptor(4, S.Dict(dict(full))),
                                          StateSpaceDescriptor(4, S.Dict({"position": xy, "size": size, "confidence": conf})),
                                          StateSpaceDescriptor(6, S.Dict({"position": xy, "confidence": conf}))],
                          targetDefs=[PredictionDescriptor(numContinuous=4, contIdx=[2, 3, 4, 5]),
                                      PredictionDescriptor(numContinuous=4, contIdx=[2, 3, 4, 5]),
                                      PredictionDescriptor(numContinuous=2, contIdx=[2, 3]), PredictionDescriptor(numContinuous=0)])


def validate_robocup_actions(actions, nAgents, allowHeadTurn):
    """The checks of RoboCupEnvironment.step / processAction (:461-463, :543-550), same messages."""
    actions = np.asarray(actions)
    if actions.shape != (nAgents, 4):
        raise Exception("Error: There must be %d actions of length 4" % nAgents)
    a = actions.astype(np.float64)
    move, turn, kick, head = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    if ((move < 0) | (move > 4)).any() or ((turn < 0) | (turn > 2)).any() or ((kick < 0) | (kick > 2)).any():
        raise Exception("Error: Robot motion commands must be between 0 and 4, 0 and 2 and 0 and 2")
    if allowHeadTurn:
        if (np.abs(head) > 3).any():
            raise Exception("Error: Head turn commands must be between -3 and +3")
    elif ((head < 0) | (head > 6)).any():
        raise Exception("Error: Head turn commands must be between 0 and 6")
    return a.astype(np.float32)


def robocup_spaces(observationType, allowHeadTurn, mean=2.0):
    """action_space / observation_space as the reference builds them (:338-430)"""
    Box, Dict, Tuple, MultiBinary, MultiDiscrete = S.Box, S.Dict, S.Tuple, S.MultiBinary, S.MultiDiscrete
    action = Tuple((MultiDiscrete([5, 3, 3]), Box(low=-3, high=3, shape=(1,)))) if allowHeadTurn else \
        Tuple((MultiDiscrete([5, 3, 3, 7]),))
    pos_xy = Box(-mean * 2, +mean * 2, shape=(2,))
    pos_radial = Box(-1, +1, shape=(3,))
    team = Box(-1, 1, shape=(1,))
    radius = Box(-mean * 2, +mean * 2, shape=(1,))
    line = Dict({"position": pos_radial, "type": Box(-1, +1, shape=(2,))})
    cross = Dict({"position": pos_radial, "radius": radius, "type": Box(-1, +1, shape=(2,))})
    fcross = Dict({"position": pos_radial, "radius": radius, "type": Box(-1, +1, shape=(2,)), "angle": Box(-1, +1, shape=(2,))})
    self_space = Dict({"position": pos_xy, "orientation": Box(-1, 1, shape=(4,)), "team": team,
                       "penalized or penalized": MultiBinary(1)})
    if observationType == ObservationType.FULL:
        ball = Dict({"position": pos_xy, "team": team, "closest": MultiBinary(1)})
        robot = Dict({"position": pos_xy, "orientation": Box(-1, 1, shape=(2,)), "team": team,
                      "penalized or penalized": MultiBinary(1)})
        obs = Tuple([Tuple([ball, robot]), Tuple([self_space, ])])
    elif observationType == ObservationType.IMAGE:
        obs = Box(0, 1, shape=(4, 480, 640))
    else:
        ball = Dict({"position": pos_xy, "radius": radius, "team": team, "closest": MultiBinary(1)})
        robot = Dict({"position": pos_xy, "radius": radius, "orientation": Box(-1, 1, shape=(2,)), "team": team,
                      "penalized or penalized": MultiBinary(1)})
        obs = Tuple([Tuple([ball, robot]), Tuple([cross, cross, fcross, line])])
    return action, obs


class _Agent(object):
    """what callers read from `env.agents[i]` (Robot.py:58-88): identity only; the state lives on the device"""

    def __init__(self, idx, nPlayers):
        self.id = idx
        self.team = 1 if idx < nPlayers else -1


class RoboCupEnvironment(object):
    """One RoboCup environment with the reference's interface, simulated on the GPU (batch of one).
    Class-level toggles as in RoboCupEnvironment.py:18-21."""
    randomInit = False
    deterministicTurn = False
    canFall = True
    useObsRewards = True

    def __init__(self, nPlayers, render=False, observationType=ObservationType.PARTIAL, noiseType=NoiseType.REALISTIC,
                 noiseMagnitude=2, obs_space_cast=False, allowHeadTurn=False, seed=0, device=0):
        if noiseMagnitude < 0 or noiseMagnitude > 5:  # environment_base.py:162-164 (which prints and exits)
            raise ValueError("Error: The noise magnitude must be between 0 and 5!")
        from .batched import BatchedRoboCup
        self.observationType = ObservationType(observationType)
        self.noiseType = NoiseType(noiseType)
        self.noiseMagnitude = noiseMagnitude
        self.allowHeadTurn = allowHeadTurn
        self.obs_space_cast = obs_space_cast
        self.maxPlayers = 5
        self.nPlayers = min(nPlayers, self.maxPlayers)
        self.nTimeSteps, self.stepIterCnt, self.maxTime = L.RC_NTICK, L.RC_SUBSTEPS, 12000
        self.stepNum = self.maxTime / self.stepIterCnt
        self.W, self.H = 1040, 740
        self.mean = 2.0  # `2.0 if ObservationType.PARTIAL else 1.0` is always 2.0 (RoboCupEnvironment.py:67)
        # render=True: renderMode 'memory' -- one top-down frame per step() / reset() in self.screenShots (the reference
        # draws every substep into a pygame window, environment_base.py:152-159; here frames come from dynenv_b200.render)
        self.renderMode, self.agentVisID, self.screenShots = ('memory' if render else False), None, []
        # agentVisID = an agent's id: its agent-view debug images (RoboCupEnvironment.py:1469-1516) of the last step's
        # ticks in self.obsVis -- IMAGE mode only, see agent_view()
        self.obsVis = []
        self.elapsed = 0
        self.agents = [_Agent(i, self.nPlayers) for i in range(2 * self.nPlayers)]
        self.action_space, self.observation_space = robocup_spaces(self.observationType, allowHeadTurn, self.mean)
        self.recoDescriptor = robocup_reco_descriptor()
        self._sim = BatchedRoboCup(1, nPlayers, self.observationType, self.noiseType, noiseMagnitude, allowHeadTurn,
                                   randomInit=type(self).randomInit, deterministicTurn=type(self).deterministicTurn,
                                   canFall=type(self).canFall, useObsRewards=type(self).useObsRewards, seed=seed, device=device)
        self._full = self.observationType == ObservationType.FULL
        self._image = self.observationType == ObservationType.IMAGE
        self._last = None
        self.reset()

    def _observations(self, out):
        """the reference's nested observation lists; IMAGE: one (8, 480, 640) float64 array per agent and tick"""
        if self._image:
            return RF.image_observations(self._sim.render_images().cpu().numpy(), 0, 2 * self.nPlayers)
        return RF.observations(self._last, 0, 2 * self.nPlayers, self._full)

    # ---- the reference's calls
    def set_random_seed(self, seed):
        """seeds the generator; like the reference (environment_base.py:190-193) it does not reset"""
        self._sim.set_random_seed(seed)

    setRandomSeed = set_random_seed  # README.md:139 spelling

    def reset(self):
        out = self._sim.reset()
        self._last = RF.host_views(out, self._full)
        self.elapsed = 0
        if self.renderMode:
            self.render()
        return self._observations(out)

    def step(self, actions):
        import torch
        a = validate_robocup_actions(actions, 2 * self.nPlayers, self.allowHeadTurn)
        out = self._sim.step(torch.from_numpy(a[None]).to(self._sim.device))
        v = self._last = RF.host_views(out, self._full)
        P = 2 * self.nPlayers
        self.elapsed += self.stepIterCnt
        if self.renderMode:
            self.render()
        # (the reference's own step() raises TypeError in IMAGE mode, RoboCupEnvironment.py:1571; here it works and
        # the observation reward is the PARTIAL one)
        return self._observations(out), v["rewards"][0, :P].astype(np.float64), bool(v["done"][0]), RF.info_dict(v, 0, P)

    def getFullState(self, robot=None):
        v, P = self._last, 2 * self.nPlayers
        if robot is None:
            return [np.array(v["state_robots"][0, :P]), np.array(v["state_ball"][0])]
        a = robot.id if hasattr(robot, "id") else int(robot)
        return (np.array(v["recon_ball"][0, a][None]), np.array(v["recon_self"][0, a][None]), np.array(v["recon_others"][0, a, :P - 1]))

    def get_agent_locs(self):
        return [self.getFullState(agent)[1][:, [0, 1, 2, 3, 4, 5]] for agent in self.agents]

    def get_full_obs(self):
        return [([o[0], o[2]], [o[1], ], (1, 1, 1)) for o in (self.getFullState(r) for r in self.agents)]

    def get_class_specific_args(self):
        return [self.allowHeadTurn]

    def render(self):
        """top-down RGB frame of the current state (uint8 [H, W, 3]); appended to screenShots in renderMode 'memory'"""
        from . import render as R
        frame = R.render_env(self._sim, 0)
        if self.renderMode == 'memory':
            self.screenShots.append(frame)
            if self.agentVisID is not None and self._image:
                self.obsVis = [self.agent_view(self.agentVisID, t) for t in range(self.nTimeSteps)]
        return frame

    def agent_view(self, agent, tick=-1):
        """the reference's agent-view debug image (RoboCupEnvironment.py:1469-1516) of one agent at one observation tick of
        the last step() / reset(): uint8 [940, 1040, 3].  Drawn on the host from the raw detection record of that (tick,
        agent), which the observation kernel emits in IMAGE mode only -- in PARTIAL mode the lists leave the device encoded
        (line rows keep distance and direction, not the end points), so there is nothing to draw from."""
        from . import render as R
        if not self._image:
            raise RuntimeError("agent_view needs observationType=ObservationType.IMAGE (raw detection records)")
        a = agent.id if hasattr(agent, "id") else int(agent)
        return R.render_agent_view(self._sim.out.raw[0, tick, a].cpu().numpy())

    def close(self):
        self._sim.close()


def validate_driving_actions(actions, nAgents, continuousActions):
    """The checks of DrivingEnvironment.step / processAction (:261-263, :364-367), same messages.  With
    continuousActions the reference itself raises UnboundLocalError (:360-368); here acc = action[0], steer = action[1]."""
    actions = np.asarray(actions)
    if actions.shape != (nAgents, 2):
        raise Exception("Error: There must be 2 actions for every car")
    a = actions.astype(np.float64)
    acc, steer = (a[:, 0], a[:, 1]) if continuousActions else (a[:, 0] - 1, (a[:, 1] - 1) * 2)
    if (np.abs(acc) > 3).any():
        raise Exception("Error: Acceleration must be between +/-3")
    if (np.abs(steer) > 3).any():
        raise Exception("Error: Steering must be between +/-3")
    return a.astype(np.float32)


def driving_spaces(observationType, continuousActions, mean=5.0):
    """action_space / observation_space as the reference builds them (DrivingEnvironment.py:170-230)"""
    Box, Dict, Tuple, MultiBinary, MultiDiscrete = S.Box, S.Dict, S.Tuple, S.MultiBinary, S.MultiDiscrete
    action = Tuple((Box(-3, +3, shape=(2,)),)) if continuousActions else Tuple((MultiDiscrete([3, 3]),))
    pos_xy = Box(-mean * 2, +mean * 2, shape=(2,))
    width_height = Box(-10, 10, shape=(2,))
    orientation = Box(-1, 1, shape=(2,))
    type_ = Box(-1, 1, shape=(1,))
    self_space = Dict({"position": pos_xy, "orientation": orientation, "width_height": width_height, "goal_position": pos_xy,
                       "finished": MultiBinary(1)})
    car_space = Dict({"position": pos_xy, "orientation": orientation, "width_height": width_height, "finished": MultiBinary(1)})
    pedestrian_space = Dict({"position": pos_xy})
    if observationType == ObservationType.FULL:
        lane_space = Dict({"points": Box(-mean * 2, mean * 2, shape=(4,)), "type": type_})
        obstacle_space = Dict({"position": pos_xy, "width_height": width_height})
    else:
        lane_space = Dict({"signed_distance": Box(-mean * 2, mean * 2, shape=(1,)), "orientation": orientation, "type": type_})
        obstacle_space = Dict({"position": pos_xy, "orientation": orientation, "width_height": width_height})
    return action, Tuple([Tuple([car_space, obstacle_space, pedestrian_space]), Tuple([self_space, lane_space])])


class DrivingEnvironment(object):
    """One Driving environment with the reference's interface (DrivingEnvironment.py:20-56), simulated on the GPU
    (batch of one)."""

    def __init__(self, nPlayers, render=False, observationType=ObservationType.PARTIAL, noiseType=NoiseType.REALISTIC,
                 noiseMagnitude=2, obs_space_cast=False, continuousActions=False, seed=0, device=0):
        if noiseMagnitude < 0 or noiseMagnitude > 5:  # environment_base.py:162-164 (which prints and exits)
            raise ValueError("Error: The noise magnitude must be between 0 and 5!")
        if ObservationType(observationType) == ObservationType.IMAGE:  # DrivingEnvironment.py:33-35 (prints and exits)
            raise ValueError("Image observation type is not supported for this environment")
        from .batched import BatchedDriving
        self.observationType = ObservationType(observationType)
        self.noiseType = NoiseType(noiseType)
        self.noiseMagnitude = noiseMagnitude
        self.continuousActions = continuousActions
        self.obs_space_cast = obs_space_cast
        self.maxPlayers = 10
        self.nPlayers = min(nPlayers, self.maxPlayers)
        self.nTimeSteps, self.stepIterCnt, self.maxTime = 1, L.DRV_SUBSTEPS, 6000
        self.stepNum = self.maxTime / self.stepIterCnt
        self.W, self.H = 1750, 1000
        self.mean = 5.0
        # render=True: renderMode 'memory' -- one top-down frame per step() / reset() in self.screenShots (the reference
        # draws every substep into a pygame window, environment_base.py:152-159; here frames come from dynenv_b200.render)
        self.renderMode, self.agentVisID, self.screenShots = ('memory' if render else False), None, []
        self.elapsed = 0
        self.agents = [_Agent(i, self.nPlayers) for i in range(self.nPlayers)]
        self.action_space, self.observation_space = driving_spaces(self.observationType, continuousActions, self.mean)
        self.recoDescriptor = driving_reco_descriptor(self.mean)
        self._sim = BatchedDriving(1, nPlayers, self.observationType, self.noiseType, noiseMagnitude, continuousActions,
                                   seed=seed, device=device)
        self._full = self.observationType == ObservationType.FULL
        self._last = None
        self.reset()

    def set_random_seed(self, seed):
        """seeds the generator; like the reference (environment_base.py:190-193) it does not reset"""
        self._sim.set_random_seed(seed)

    setRandomSeed = set_random_seed  # README.md:139 spelling

    def reset(self):
        self._last = RF.drv_host_views(self._sim.reset())
        self.elapsed = 0
        if self.renderMode:
            self.render()
        return RF.drv_observations(self._last, 0, self.nPlayers, self._full)

    def step(self, actions):
        import torch
        a = validate_driving_actions(actions, self.nPlayers, self.continuousActions)
        out = self._sim.step(torch.from_numpy(a[None]).to(self._sim.device))
        v = self._last = RF.drv_host_views(out)
        N = self.nPlayers
        self.elapsed += self.stepIterCnt
        if self.renderMode:
            self.render()
        return RF.drv_observations(v, 0, N, self._full), v["rewards"][0, :N].astype(np.float64), bool(v["done"][0]), \
            RF.drv_info_dict(v, 0, N)

    def getFullState(self, agent=None):
        a = None if agent is None else (agent.id if hasattr(agent, "id") else int(agent))
        return RF.drv_full_state(self._last, 0, self.nPlayers, a)

    def get_agent_locs(self):
        return [self.getFullState(agent)[0][:, [0, 1, 2, 3]] for agent in self.agents]

    def get_full_obs(self):
        return [([o[1], o[2], o[3]], [o[0], o[4]], (1, 1, 1)) for o in (self.getFullState(car) for car in self.agents)]

    def get_class_specific_args(self):
        return [self.continuousActions]

    def render(self):
        """top-down RGB frame of the current state (uint8 [H, W, 3]); appended to screenShots in renderMode 'memory'"""
        from . import render as R
        frame = R.render_env(self._sim, 0)
        if self.renderMode == 'memory':
            self.screenShots.append(frame)
        return frame

    def close(self):
        self._sim.close()


class BatchedVecEnv(object):
    """VecEnv over one batched simulator: `num_envs` environments, auto-reset on done with
    info['terminal_observation'] (utils/subproc_vec_env.py:17-23).  `step()` returns the reference's host
    format (object ndarray [nEnvs, nTime, nAgents, 3], rewards [nEnvs, nAgents], dones, infos); at large
    batch sizes use `step_tensors()`, which leaves everything on the device."""

    def __init__(self, num_envs, nPlayers, observationType, noiseType, noiseMagnitude, allowHeadTurn=False, seed=0,
                 device=0, env_id_offset=0, env=DynEnvType.ROBO_CUP, render=False, **toggles):
        from .batched import BatchedDriving, BatchedRoboCup
        self.num_envs = int(num_envs)
        # render=True (make_dyn_env passes it through, DynEnv/__init__.py:9-13): renderMode 'memory' -- one top-down frame
        # of environment 0 per step() / reset() in self.screenShots (dynenv_b200.render; the reference opens a pygame window)
        self.renderMode, self.screenShots = ('memory' if render else False), []
        self.driving = env is DynEnvType.DRIVE
        self.allowHeadTurn = allowHeadTurn  # RoboCup: allowHeadTurn, Driving: continuousActions (make_dyn_env)
        self.observationType = ObservationType(observationType)
        self._full = self.observationType == ObservationType.FULL
        if self.driving:
            self.nPlayers = self.nAgents = min(nPlayers, 10)
            self.nAct = 2
            self.action_space, self.observation_space = driving_spaces(self.observationType, allowHeadTurn)
            self.sim = BatchedDriving(self.num_envs, nPlayers, observationType, noiseType, noiseMagnitude, allowHeadTurn,
                                      auto_reset=True, terminal_outputs=True, seed=seed, device=device,
                                      env_id_offset=env_id_offset)
            nT, nSub, maxT = 1, L.DRV_SUBSTEPS, 6000
        else:
            self.nPlayers = min(nPlayers, 5)
            self.nAgents = 2 * self.nPlayers
            self.nAct = 4
            self.action_space, self.observation_space = robocup_spaces(self.observationType, allowHeadTurn)
            self.sim = BatchedRoboCup(self.num_envs, nPlayers, observationType, noiseType, noiseMagnitude, allowHeadTurn,
                                      auto_reset=True, terminal_outputs=True, seed=seed, device=device,
                                      env_id_offset=env_id_offset, **toggles)
            nT, nSub, maxT = L.RC_NTICK, L.RC_SUBSTEPS, 12000
        self._actions = None
        self.recoDescriptor = driving_reco_descriptor() if self.driving else robocup_reco_descriptor()
        self._attrs = dict(nPlayers=self.nPlayers, nTimeSteps=nT, stepIterCnt=nSub, maxTime=maxT,
                           stepNum=maxT / nSub, observationType=self.observationType, renderMode=self.renderMode,
                           agentVisID=None, action_space=self.action_space, observation_space=self.observation_space,
                           recoDescriptor=self.recoDescriptor)

    # ---- format adapters (RoboCup / Driving)
    def _views(self, out, env_ids=None):
        if self.driving:
            return RF.drv_host_views(out, env_ids)
        v = RF.host_views(out, self._full, env_ids)
        if self.observationType == ObservationType.IMAGE:  # host format: every image is materialised (small batches)
            imgs = self.sim.render_images(outputs=out).cpu().numpy()
            v["images"] = imgs if env_ids is None else imgs[env_ids]
        return v

    def _obs(self, v, e):
        if self.driving:
            return RF.drv_observations(v, e, self.nAgents, self._full)
        if "images" in v:  # IMAGE mode: v["images"] holds the rendered planes of the environments in v
            return RF.image_observations(v["images"], e, self.nAgents)
        return RF.observations(v, e, self.nAgents, self._full)

    def _info(self, v, e):
        return RF.drv_info_dict(v, e, self.nAgents) if self.driving else RF.info_dict(v, e, self.nAgents)

    def _validate(self, a):
        if self.driving:
            return validate_driving_actions(a, self.nAgents, self.allowHeadTurn)
        return validate_robocup_actions(a, self.nAgents, self.allowHeadTurn)

    # ---- device-resident fast path
    def step_tensors(self, actions):
        """actions: float32 CUDA tensor [num_envs, nAgents, 4]; returns (out, terminal) RoboCupOutputs views"""
        return self.sim.step(actions), self.sim.terminal

    # ---- VecEnv calls
    def render(self, env=0):
        """top-down RGB frame of one environment (uint8 [H, W, 3]); appended to screenShots in renderMode 'memory'"""
        from . import render as R
        frame = R.render_env(self.sim, env)
        if self.renderMode == 'memory':
            self.screenShots.append(frame)
        return frame

    def reset(self):
        v = self._views(self.sim.reset())
        if self.renderMode:
            self.render()
        return self._stack([self._obs(v, e) for e in range(self.num_envs)])

    def step_async(self, actions):
        self._actions = actions

    def step_wait(self):
        import torch
        a = np.asarray(self._actions)
        if a.shape != (self.num_envs, self.nAgents, self.nAct):
            raise Exception("Error: There must be 2 actions for every car" if self.driving else
                            "Error: There must be %d actions of length 4" % self.nAgents)
        a = np.stack([self._validate(a[e]) for e in range(self.num_envs)])
        out = self.sim.step(torch.from_numpy(a).to(self.sim.device))
        v = self._views(out)
        dones = v["done"].astype(bool)
        tv = self._views(self.sim.terminal, np.nonzero(dones)[0]) if dones.any() else None
        obs, infos, k = [], [], 0
        for e in range(self.num_envs):
            info = self._info(v, e)
            if dones[e]:
                # obs/full of the finished step were routed to the terminal buffers by the kernel
                info["terminal_observation"] = self._obs(dict(tv), k)
                k += 1
            obs.append(self._obs(v, e))
            infos.append(info)
        if self.renderMode:
            self.render()
        return self._stack(obs), v["rewards"][:, :self.nAgents].astype(np.float64), dones, tuple(infos)

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    @staticmethod
    def _stack(obs):
        """_flatten_obs (subproc_vec_env.py:178-201): object array [nEnvs, nTime, nAgents, 3]"""
        nE, nT, nA = len(obs), len(obs[0]), len(obs[0][0])
        arr = np.empty((nE, nT, nA, 3), dtype=object)
        for e in range(nE):
            for t in range(nT):
                for a in range(nA):
                    for k in range(3):
                        arr[e, t, a, k] = obs[e][t][a][k]
        return arr

    def seed(self, seed=None):
        if seed is not None:
            self.sim.set_random_seed(seed)  # global env ids keep the per-env streams distinct

    def get_attr(self, attr_name, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [self._attrs[attr_name]] * n

    def env_method(self, method_name, *args, **kwargs):
        if method_name == "get_agent_locs" and self.driving:  # DrivingEnvironment.py:126-127
            sc = self.sim.out.state_cars.detach().cpu().numpy()
            return [[sc[e, a][None][:, [0, 1, 2, 3]] for a in range(self.nAgents)] for e in range(self.num_envs)]
        if method_name == "get_agent_locs":
            rs = self.sim.out.recon_self.detach().cpu().numpy()
            return [[rs[e, a][None][:, [0, 1, 2, 3, 4, 5]] for a in range(self.nAgents)] for e in range(self.num_envs)]
        raise AttributeError(method_name)

    def close(self):
        self.sim.close()


def make_dyn_env(env, num_envs, num_players, render, observationType, noiseType, noiseMagnitude,
                 use_continuous_actions, seed=0, device=0):
    """DynEnv.make_dyn_env (DynEnv/__init__.py:6-25): same arguments, same return (vec_env, name).  render=True keeps
    top-down frames of environment 0 in vec_env.screenShots (the reference's environments open a pygame window each)."""
    if env is DynEnvType.ROBO_CUP:
        return BatchedVecEnv(num_envs, num_players, observationType, noiseType, noiseMagnitude, render=bool(render),
                             allowHeadTurn=use_continuous_actions, seed=seed, device=device), "RoboCup"
    if env is DynEnvType.DRIVE:
        return BatchedVecEnv(num_envs, num_players, observationType, noiseType, noiseMagnitude, render=bool(render),
                             allowHeadTurn=use_continuous_actions, seed=seed, device=device, env=DynEnvType.DRIVE), "Driving"
    raise ValueError
