"""Problem descriptors for the reference examples on the hot path.

Each builder restates what the corresponding `Problem` subclass of the reference
constructs (dynamics, player costs, initial state) as the POD `ilqg_problem_desc`
of include/ilqg.h.  Record order per player follows the reference's accumulation
order (src/player_cost.cpp:194-215): state costs, control costs, state
constraints, control constraints.

  three_player_intersection  src/three_player_intersection_example.cpp:74-394
  roundabout_merging         src/roundabout_merging_example.cpp:71-436,
                             src/roundabout_lane_center.cpp:50-108
  air_3d                     src/air_3d_example.cpp:62-141
  two_player_point_mass_lq   test/test_lq_solver.cpp:143-177,227-248 (LQ-only)
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

from . import _abi as abi

F = np.float32


class DescBuilder:
    def __init__(self, num_time_steps: int = 100, time_step: float = 0.1):
        self.d = abi.ProblemDesc()
        self.d.num_time_steps = num_time_steps
        self.d.time_step = time_step
        self.d.initial_time = 0.0
        self.d.num_players = 0
        self.d.xdim = 0
        self.d.num_subsystems = 0
        self.d.num_costs = 0
        self.d.num_polylines = 0
        self.d.polyline_start[0] = 0
        self._records: List[List[List[abi.CostDesc]]] = []  # [player][category] -> records

    # dynamics ---------------------------------------------------------------
    def add_player(self, udim: int, state_reg: float = 0.0, control_reg: float = 0.0,
                   structure: int = abi.COST_SUM) -> int:
        i = self.d.num_players
        self.d.udim[i] = udim
        self.d.state_regularization[i] = state_reg
        self.d.control_regularization[i] = control_reg
        self.d.cost_structure[i] = structure
        self.d.num_players += 1
        self._records.append([[], [], [], []])
        return i

    def add_subsystem(self, kind: int, xdim: int, first_player: int, params: Sequence[float] = ()):
        s = self.d.subsystems[self.d.num_subsystems]
        s.kind = kind
        s.x_offset = self.d.xdim
        s.first_player = first_player
        for k, p in enumerate(params):
            s.params[k] = p
        self.d.xdim += xdim
        self.d.num_subsystems += 1
        return s.x_offset

    # geometry ---------------------------------------------------------------
    def add_polyline(self, points: Sequence[Tuple[float, float]]) -> int:
        p = self.d.num_polylines
        start = self.d.polyline_start[p]
        assert len(points) >= 2 and start + len(points) <= abi.MAX_POLYLINE_POINTS
        for k, (x, y) in enumerate(points):
            self.d.polyline_points[start + k][0] = float(F(x))
            self.d.polyline_points[start + k][1] = float(F(y))
        self.d.polyline_start[p + 1] = start + len(points)
        self.d.num_polylines += 1
        return p

    # costs ------------------------------------------------------------------
    def _rec(self, player, category, kind, arg=-1, dims=(0,), weight=1.0, value=0.0, flag=0,
             polyline=-1, is_equality=0, active_from=0.0, group=0, group_is_min=0):
        r = abi.CostDesc()
        r.active_from = active_from   # FinalTimeCost threshold (0 = always)
        r.group, r.group_is_min = group, int(group_is_min)   # ExtremeValueCost membership (0 = none)
        r.kind, r.player, r.arg, r.is_equality = kind, player, arg, is_equality
        for k in range(4):
            r.dim[k] = dims[k] if k < len(dims) else 0
        r.flag, r.polyline, r.weight, r.value = int(flag), polyline, weight, value
        self._records[player][category].append(r)

    def state_cost(self, player, kind, **kw):
        self._rec(player, 0, kind, arg=-1, **kw)

    def control_cost(self, player, control_of, kind, **kw):
        self._rec(player, 1, kind, arg=control_of, **kw)

    def state_constraint(self, player, kind, **kw):
        self._rec(player, 2, kind, arg=-1, **kw)

    def control_constraint(self, player, control_of, kind, **kw):
        self._rec(player, 3, kind, arg=control_of, **kw)

    def build(self) -> abi.ProblemDesc:
        c = 0
        for player_records in self._records:
            for category in player_records:
                for r in category:
                    assert c < abi.MAX_COSTS
                    self.d.costs[c] = r
                    c += 1
        self.d.num_costs = c
        return self.d


# --------------------------------------------------------------------------
# ThreePlayerIntersectionExample, src/three_player_intersection_example.cpp
# --------------------------------------------------------------------------
def three_player_intersection(num_time_steps: int = 100, time_step: float = 0.1,
                              with_constraints: bool = True, lane_constraints: bool = False):
    """Returns (desc, x0).  2x SinglePlayerCar6D + 1x SinglePlayerUnicycle4D, n = 16.
    with_constraints = False drops the proximity constraints (an unconstrained variant used by the
    receding-horizon tests; the reference example always has them).
    lane_constraints = True adds the six Polyline2SignedDistanceConstraints the example constructs
    and leaves commented out (its lane boundaries, :214-251), in the order the commented lines would
    add them: each player's right and left boundary ahead of its proximity constraints."""
    b = DescBuilder(num_time_steps, time_step)
    kInterAxleLength = 4.0
    kStateReg, kControlReg = 1.0, 5.0
    kOmegaCostWeight, kJerkCostWeight, kACostWeight = 0.1, 0.1, 0.1
    kNominalVCostWeight, kLaneCostWeight, kMinProximity = 100.0, 25.0, 6.0
    kLaneHalfWidth, kOrientedRight = 2.5, True  # :99,103
    kP1NominalV, kP2NominalV, kP3NominalV = 8.0, 5.0, 1.5
    kP1InitialX, kP2InitialX, kP3InitialX = -2.0, -10.0, -11.0
    kP1InitialY, kP2InitialY, kP3InitialY = -30.0, 45.0, 16.0
    kP1InitialHeading, kP2InitialHeading, kP3InitialHeading = math.pi / 2, -math.pi / 2, 0.0
    kP1InitialSpeed, kP2InitialSpeed, kP3InitialSpeed = 4.0, 3.0, 1.25

    for _ in range(3):  # :191-196
        b.add_player(2, kStateReg, kControlReg)
    o1 = b.add_subsystem(abi.DYN_CAR6D, 6, 0, [kInterAxleLength])  # :166-170
    o2 = b.add_subsystem(abi.DYN_CAR6D, 6, 1, [kInterAxleLength])
    o3 = b.add_subsystem(abi.DYN_UNICYCLE4D, 4, 2)
    # index constants :133-163
    kP1XIdx, kP1YIdx, kP1HeadingIdx, kP1VIdx = o1 + 0, o1 + 1, o1 + 2, o1 + 4
    kP2XIdx, kP2YIdx, kP2HeadingIdx, kP2VIdx = o2 + 0, o2 + 1, o2 + 2, o2 + 4
    kP3XIdx, kP3YIdx, kP3HeadingIdx, kP3VIdx = o3 + 0, o3 + 1, o3 + 2, o3 + 3

    lane1 = b.add_polyline([(kP1InitialX, -1000.0), (kP1InitialX, 1000.0)])  # :201-209
    lane2 = b.add_polyline([(kP2InitialX, 1000.0), (kP2InitialX, 18.0), (kP2InitialX + 0.5, 15.0),
                            (kP2InitialX + 1.0, 14.0), (kP2InitialX + 3.0, 12.5),
                            (kP2InitialX + 6.0, 12.0), (1000.0, 12.0)])
    lane3 = b.add_polyline([(-1000.0, kP3InitialY), (1000.0, kP3InitialY)])

    pos = [(kP1XIdx, kP1YIdx), (kP2XIdx, kP2YIdx), (kP3XIdx, kP3YIdx)]
    vidx = [kP1VIdx, kP2VIdx, kP3VIdx]
    nominal_v = [kP1NominalV, kP2NominalV, kP3NominalV]
    lanes = [lane1, lane2, lane3]
    for i in range(3):
        b.state_cost(i, abi.COST_QUADRATIC_POLYLINE2, dims=pos[i], weight=kLaneCostWeight,
                     polyline=lanes[i])  # :211-251
        b.state_cost(i, abi.COST_QUADRATIC, dims=(vidx[i],), weight=kNominalVCostWeight,
                     value=nominal_v[i])  # :254-287
        if lane_constraints:  # :214-221, 229-236, 244-251 (commented out in the reference example)
            b.state_constraint(i, abi.CONSTRAINT_POLYLINE2_SIGNED_DISTANCE, dims=pos[i], polyline=lanes[i],
                               value=kLaneHalfWidth, flag=int(not kOrientedRight))
            b.state_constraint(i, abi.CONSTRAINT_POLYLINE2_SIGNED_DISTANCE, dims=pos[i], polyline=lanes[i],
                               value=-kLaneHalfWidth, flag=int(kOrientedRight))
    # control costs :289-335 (P3's second control is acceleration)
    b.control_cost(0, 0, abi.COST_QUADRATIC, dims=(0,), weight=kOmegaCostWeight, value=0.0)
    b.control_cost(0, 0, abi.COST_QUADRATIC, dims=(1,), weight=kJerkCostWeight, value=0.0)
    b.control_cost(1, 1, abi.COST_QUADRATIC, dims=(0,), weight=kOmegaCostWeight, value=0.0)
    b.control_cost(1, 1, abi.COST_QUADRATIC, dims=(1,), weight=kJerkCostWeight, value=0.0)
    b.control_cost(2, 2, abi.COST_QUADRATIC, dims=(0,), weight=kOmegaCostWeight, value=0.0)
    b.control_cost(2, 2, abi.COST_QUADRATIC, dims=(1,), weight=kACostWeight, value=0.0)
    # collision-avoidance constraints :363-394 (keep_within = !kKeepClose = false)
    for i, others in ((0, (1, 2)), (1, (0, 2)), (2, (0, 1))) if with_constraints else ():
        for j in others:
            b.state_constraint(i, abi.CONSTRAINT_PROXIMITY, dims=pos[i] + pos[j],
                               value=kMinProximity, flag=0)

    x0 = np.zeros(b.d.xdim, dtype=F)  # :172-185
    x0[kP1XIdx], x0[kP1YIdx], x0[kP1HeadingIdx], x0[kP1VIdx] = (
        kP1InitialX, kP1InitialY, kP1InitialHeading, kP1InitialSpeed)
    x0[kP2XIdx], x0[kP2YIdx], x0[kP2HeadingIdx], x0[kP2VIdx] = (
        kP2InitialX, kP2InitialY, kP2InitialHeading, kP2InitialSpeed)
    x0[kP3XIdx], x0[kP3YIdx], x0[kP3HeadingIdx], x0[kP3VIdx] = (
        kP3InitialX, kP3InitialY, kP3InitialHeading, kP3InitialSpeed)
    return b.build(), x0


def three_player_intersection_params(**overrides) -> abi.SolverParams:
    """SolverParams as set by exec/three_player_intersection/main.cpp:109-120."""
    base = dict(max_backtracking_steps=100, max_solver_iters=100,
                unconstrained_solver_max_iters=10, linesearch=1,
                expected_decrease_fraction=0.001, initial_alpha_scaling=0.1,
                convergence_tolerance=1.0, geometric_mu_scaling=1.1,
                geometric_mu_downscaling=0.5, geometric_lambda_downscaling=0.5)
    base.update(overrides)
    return abi.SolverParams.defaults(**base)


def three_player_intersection_x0_batch(batch: int, seed: int) -> np.ndarray:
    """SURVEY section 8d input #2: example x0 + positions U(-2,2) m, headings
    U(-0.1,0.1) rad, speeds x U(0.8,1.2)."""
    _, x0 = three_player_intersection()
    rng = np.random.default_rng(seed)
    out = np.tile(x0, (batch, 1)).astype(F)
    pos = [0, 1, 6, 7, 12, 13]
    head = [2, 8, 14]
    spd = [4, 10, 15]
    out[:, pos] += rng.uniform(-2.0, 2.0, size=(batch, len(pos))).astype(F)
    out[:, head] += rng.uniform(-0.1, 0.1, size=(batch, len(head))).astype(F)
    out[:, spd] *= rng.uniform(0.8, 1.2, size=(batch, len(spd))).astype(F)
    return out


# --------------------------------------------------------------------------
# RoundaboutMergingExample
# --------------------------------------------------------------------------
def _libm():
    """The C library's float math: std::cos(float) etc. in the reference are cosf/sinf/atan2f, which
    are not correctly rounded, so rounding a double result to fp32 differs from them in the last
    bit now and then (caught by tests/golden/make_ref_golden.py:check_polylines)."""
    import ctypes
    import ctypes.util
    global _LIBM
    if _LIBM is None:
        _LIBM = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        for name, nargs in (("cosf", 1), ("sinf", 1), ("atan2f", 2)):
            fn = getattr(_LIBM, name)
            fn.restype = ctypes.c_float
            fn.argtypes = [ctypes.c_float] * nargs
    return _LIBM


_LIBM = None


def _cosf(a) -> np.float32:
    return F(_libm().cosf(float(F(a))))


def _sinf(a) -> np.float32:
    return F(_libm().sinf(float(F(a))))


def _atan2f(y, x) -> np.float32:
    return F(_libm().atan2f(float(F(y)), float(F(x))))


def roundabout_lane_center(entrance_angle, exit_angle, distance_from_roundabout):
    """RoundaboutLaneCenter, src/roundabout_lane_center.cpp:50-108 (fp32 arithmetic)."""
    kR, kH = F(12.0), F(2.5)
    ea, xa, dist = F(entrance_angle), F(exit_angle), F(distance_from_roundabout)
    cos, sin = _cosf, _sinf
    cx, cy = (kR + kH) * cos(ea), (kR + kH) * sin(ea)
    first_angle = F(float(ea) - math.pi / 2)
    fx, fy = cx + kH * cos(first_angle), cy + kH * sin(first_angle)
    pts = [(fx + dist * cos(ea), fy + dist * sin(ea)), (fx, fy)]
    for ii in range(1, 4):
        ang = F(float(first_angle) - (math.pi / 2) * float(F(ii)) / 3)
        pts.append((cx + kH * cos(ang), cy + kH * sin(ang)))
    for ii in range(1, 11):
        nxt = F(ea + (xa - ea) * F(ii) / F(10))
        pts.append((kR * cos(nxt), kR * sin(nxt)))
    pts.append((F(1e4) * cos(xa), F(1e4) * sin(xa)))
    return [(float(x), float(y)) for x, y in pts]


def roundabout_merging(num_time_steps: int = 100, time_step: float = 0.1):
    """Returns (desc, x0).  4x SinglePlayerCar6D, n = 24, unconstrained."""
    b = DescBuilder(num_time_steps, time_step)
    kOmegaCostWeight, kACostWeight, kJerkCostWeight = 500.0, 50.0, 5.0
    kMaxVCostWeight, kNominalVCostWeight = 1000.0, 10.0
    kLaneCostWeight, kLaneBoundaryCostWeight = 25.0, 100.0
    kMinProximity, kProximityCostWeight = 6.0, 100.0
    kLaneHalfWidth, kMaxV, kMinV, kNominalV = 2.5, 12.0, 1.0, 10.0
    dist = [25.0, 10.0, 25.0, 10.0]
    speed = [3.0, 2.0, 3.0, 2.0]
    kAngleOffset = F(math.pi / 2 * 0.5)  # static constexpr float
    kWedgeSize = F(math.pi)
    angles = [F(kAngleOffset), F(float(kAngleOffset) + 2.0 * math.pi / 4.0),
              F(float(kAngleOffset) + 2.0 * 2.0 * math.pi / 4.0),
              F(float(kAngleOffset) + 3.0 * 2.0 * math.pi / 4.0)]
    offs = []
    for i in range(4):
        b.add_player(2, 0.0, 0.0)
    for i in range(4):
        offs.append(b.add_subsystem(abi.DYN_CAR6D, 6, i, [4.0]))
    lanes, lane_pts = [], []
    for i in range(4):
        pts = roundabout_lane_center(angles[i], F(angles[i] + kWedgeSize), dist[i])
        lane_pts.append(pts)
        lanes.append(b.add_polyline(pts))
    pos = [(o + 0, o + 1) for o in offs]
    vidx = [o + 4 for o in offs]
    kP1AIdx = offs[0] + 5
    for i in range(4):
        # lane costs :211-266
        b.state_cost(i, abi.COST_QUADRATIC_POLYLINE2, dims=pos[i], weight=kLaneCostWeight,
                     polyline=lanes[i])
        b.state_cost(i, abi.COST_SEMIQUADRATIC_POLYLINE2, dims=pos[i],
                     weight=kLaneBoundaryCostWeight, polyline=lanes[i], value=kLaneHalfWidth,
                     flag=1)
        b.state_cost(i, abi.COST_SEMIQUADRATIC_POLYLINE2, dims=pos[i],
                     weight=kLaneBoundaryCostWeight, polyline=lanes[i], value=-kLaneHalfWidth,
                     flag=0)
    for i in range(4):
        # speed costs :268-311
        b.state_cost(i, abi.COST_SEMIQUADRATIC, dims=(vidx[i],), weight=kMaxVCostWeight,
                     value=kMinV, flag=0)
        b.state_cost(i, abi.COST_SEMIQUADRATIC, dims=(vidx[i],), weight=kMaxVCostWeight,
                     value=kMaxV, flag=1)
        b.state_cost(i, abi.COST_QUADRATIC, dims=(vidx[i],), weight=kNominalVCostWeight,
                     value=kNominalV)
    for i in range(4):
        # acceleration cost: all four players index kP1AIdx (:342-353)
        b.state_cost(i, abi.COST_QUADRATIC, dims=(kP1AIdx,), weight=kACostWeight, value=0.0)
    for i in range(4):
        b.control_cost(i, i, abi.COST_QUADRATIC, dims=(0,), weight=kOmegaCostWeight, value=0.0)
        b.control_cost(i, i, abi.COST_QUADRATIC, dims=(1,), weight=kJerkCostWeight, value=0.0)
    # proximity costs actually added (:385-434): p1:{p2,p4} p2:{p1,p3} p3:{p2,p4} p4:{p1,p3}
    for i, others in ((0, (1, 3)), (1, (0, 2)), (2, (1, 3)), (3, (0, 2))):
        for j in others:
            b.state_cost(i, abi.COST_PROXIMITY, dims=pos[i] + pos[j],
                         weight=kProximityCostWeight, value=kMinProximity)
    # NB: the reference adds each player's state costs in source order lane(3), speed(3),
    # accel(1), proximity(2); DescBuilder keeps per-player append order, which is that order.
    x0 = np.zeros(b.d.xdim, dtype=F)  # :187-205
    for i in range(4):
        (ax, ay), (bx, by) = lane_pts[i][0], lane_pts[i][1]
        x0[offs[i] + 0] = ax
        x0[offs[i] + 1] = ay
        # LineSegment2::Heading (line_segment2.h:55-70): atan2f of the fp32 unit direction
        dx, dy = F(F(bx) - F(ax)), F(F(by) - F(ay))
        ex, ey = F(F(ax) - F(bx)), F(F(ay) - F(by))
        length = F(np.sqrt(F(F(ex * ex) + F(ey * ey))))
        x0[offs[i] + 2] = _atan2f(F(dy / length), F(dx / length))
        x0[offs[i] + 4] = speed[i]
    return b.build(), x0


def roundabout_x0_batch(batch: int, seed: int) -> np.ndarray:
    """Synthetic initial states for RoundaboutMerging: each car U(0,8) m FORWARD along its
    first lane segment (all first segments are >= 10 m long) with a lateral offset U(-1,1) m,
    heading of the segment, speed U(1,4) m/s.

    SURVEY section 8d proposed "U(0,10) m back along the first segment"; that puts every car
    exactly on the axis behind the polyline's first point, where the reference's signed
    distance is sgn(cross product ~ 0) * d^2 (src/line_segment2.cpp:56-100), i.e. decided by
    rounding noise, so fp32 implementations legitimately disagree on which lane-boundary cost
    is active.  The forward/lateral placement is well posed.
    """
    desc, x0 = roundabout_merging()
    rng = np.random.default_rng(seed)
    out = np.tile(x0, (batch, 1)).astype(F)
    fwd = rng.uniform(0.0, 8.0, size=(batch, 4)).astype(F)
    lat = rng.uniform(-1.0, 1.0, size=(batch, 4)).astype(F)
    spd = rng.uniform(1.0, 4.0, size=(batch, 4)).astype(F)
    for i in range(4):
        th = out[:, 6 * i + 2]
        out[:, 6 * i + 0] += fwd[:, i] * np.cos(th) - lat[:, i] * np.sin(th)
        out[:, 6 * i + 1] += fwd[:, i] * np.sin(th) + lat[:, i] * np.cos(th)
        out[:, 6 * i + 4] = spd[:, i]
    return out


def roundabout_params(**overrides) -> abi.SolverParams:
    """SolverParams of exec/roundabout_merging_example/main.cpp:72-75,108-113."""
    base = dict(max_backtracking_steps=100, linesearch=1, expected_decrease_fraction=0.1,
                initial_alpha_scaling=0.75, convergence_tolerance=0.01)
    base.update(overrides)
    return abi.SolverParams.defaults(**base)


# --------------------------------------------------------------------------
# ThreePlayerOvertakingExample (src/three_player_overtaking_example.cpp)
# --------------------------------------------------------------------------
def three_player_overtaking(num_time_steps: int = 100, time_step: float = 0.1):
    """Returns (desc, x0).  3x SinglePlayerCar6D on a two-lane straight road, n = 18,
    unconstrained; only record kinds the three headline examples already use."""
    b = DescBuilder(num_time_steps, time_step)
    kOmegaCostWeight, kJerkCostWeight = 500000.0, 500.0
    nominal_v_weight, nominal_v = [10.0, 1.0, 1.0], [15.0, 10.0, 10.0]
    kLaneCostWeight, kLaneBoundaryCostWeight, kLaneHalfWidth = 25.0, 100.0, 2.5
    kMinProximity, kProximityCostWeight = 5.0, 100.0
    start = [(2.5, -10.0, 10.0), (-1.0, -10.0, 2.0), (2.5, 10.0, 2.0)]   # x, y, speed (:113-130)
    for _ in range(3):
        b.add_player(2, 0.0, 0.0)                                          # PlayerCost("P1") ... (:198-200)
    offs = [b.add_subsystem(abi.DYN_CAR6D, 6, i, [4.0]) for i in range(3)]
    lane1 = b.add_polyline([(start[1][0], -1000.0), (start[1][0], 1000.0)])  # :206-209
    lane2 = b.add_polyline([(start[2][0], -1000.0), (start[2][0], 1000.0)])
    pos = [(o + 0, o + 1) for o in offs]
    for i, lane in enumerate((lane1, lane1, lane2)):                       # :211-254
        b.state_cost(i, abi.COST_QUADRATIC_POLYLINE2, dims=pos[i], weight=kLaneCostWeight, polyline=lane)
        b.state_cost(i, abi.COST_SEMIQUADRATIC_POLYLINE2, dims=pos[i], weight=kLaneBoundaryCostWeight,
                     polyline=lane, value=kLaneHalfWidth, flag=1)
        b.state_cost(i, abi.COST_SEMIQUADRATIC_POLYLINE2, dims=pos[i], weight=kLaneBoundaryCostWeight,
                     polyline=lane, value=-kLaneHalfWidth, flag=0)
    for i in range(3):                                                     # :262-286
        b.state_cost(i, abi.COST_QUADRATIC, dims=(offs[i] + 4,), weight=nominal_v_weight[i], value=nominal_v[i])
    for i in range(3):                                                     # :289-308
        b.control_cost(i, i, abi.COST_QUADRATIC, dims=(0,), weight=kOmegaCostWeight, value=0.0)
        b.control_cost(i, i, abi.COST_QUADRATIC, dims=(1,), weight=kJerkCostWeight, value=0.0)
    for i, others in ((0, (1, 2)), (1, (0, 2))):                           # :311-327; P3's are commented out
        for j in others:
            b.state_cost(i, abi.COST_PROXIMITY, dims=pos[i] + pos[j], weight=kProximityCostWeight,
                         value=kMinProximity)
    x0 = np.zeros(b.d.xdim, dtype=F)                                       # :181-195
    for i, (x, y, v) in enumerate(start):
        x0[offs[i] + 0], x0[offs[i] + 1], x0[offs[i] + 2], x0[offs[i] + 4] = x, y, F(math.pi / 2), v
    return b.build(), x0


def three_player_overtaking_params(**overrides) -> abi.SolverParams:
    """SolverParams of exec/three_player_overtaking/main.cpp:71-74,108-112."""
    base = dict(max_backtracking_steps=100, linesearch=1, expected_decrease_fraction=0.1,
                initial_alpha_scaling=0.75, convergence_tolerance=0.01)
    base.update(overrides)
    return abi.SolverParams.defaults(**base)


def three_player_overtaking_x0_batch(batch: int, seed: int) -> np.ndarray:
    """Synthetic initial states: the example's, with positions moved U(-1, 1) m, headings
    U(-0.05, 0.05) rad and speeds scaled U(0.8, 1.2)."""
    _, x0 = three_player_overtaking()
    rng = np.random.default_rng(seed)
    out = np.tile(x0, (batch, 1))
    for o in (0, 6, 12):
        out[:, o:o + 2] += rng.uniform(-1.0, 1.0, size=(batch, 2)).astype(F)
        out[:, o + 2] += rng.uniform(-0.05, 0.05, size=batch).astype(F)
        out[:, o + 4] *= rng.uniform(0.8, 1.2, size=batch).astype(F)
    return out.astype(F)


# --------------------------------------------------------------------------
# TwoPlayerCollisionExample (src/two_player_collision_example.cpp)
# --------------------------------------------------------------------------
def two_player_collision(num_time_steps: int = 100, time_step: float = 0.1):
    """Returns (desc, x0).  2x SinglePlayerCar6D driving at each other, n = 12, unconstrained; goal
    costs are FinalTimeCost-wrapped QuadraticCosts (records with active_from = 9.5 s).  CPU oracle
    only for now: the CUDA library has no time gate yet and answers ILQG_ERR_UNSUPPORTED."""
    b = DescBuilder(num_time_steps, time_step)
    kOmegaCostWeight, kJerkCostWeight = 5000.0, 3250.0
    kLaneCostWeight, kLaneBoundaryCostWeight, kLaneHalfWidth = F(250.0), F(50000.0), F(2.5)
    kMinProximity, kProximityCostWeight, kGoalCostWeight = 7.5, 5000.0, 1000.0
    for _ in range(2):
        b.add_player(2, 1.0, 0.0)                                         # PlayerCost("P1", 1.0, 0.0) (:175-176)
    offs = [b.add_subsystem(abi.DYN_CAR6D, 6, i, [4.0]) for i in range(2)]
    pos = [(o + 0, o + 1) for o in offs]
    edge = float(F(2.5 + float(kLaneHalfWidth)))
    lane1_p1p2 = b.add_polyline([(2.5, -50.0), (2.5, 50.0)])               # :182
    # player 1 (:183-245): lane centre, left boundary, five boundary pieces around the crossing
    b.state_cost(0, abi.COST_QUADRATIC_POLYLINE2, dims=pos[0], weight=kLaneCostWeight, polyline=lane1_p1p2)
    b.state_cost(0, abi.COST_SEMIQUADRATIC_POLYLINE2, dims=pos[0], weight=F(kLaneBoundaryCostWeight * F(1000)),
                 polyline=lane1_p1p2, value=-kLaneHalfWidth, flag=0)
    pieces = [([(edge, -50.0), (edge, -5.0)], 1), ([(edge, 5.0), (edge, 50.0)], 1), ([(10.0, -5.0), (10.0, 5.0)], 1),
              ([(edge, 5.0), (25.0, 5.0)], 0), ([(edge, -5.0), (25.0, -5.0)], 1)]
    # player 2 (:186-209): lane centre, left and right boundary -- declared between player 1's
    # costs in the source, but each PlayerCost keeps its own order
    b.state_cost(1, abi.COST_QUADRATIC_POLYLINE2, dims=pos[1], weight=F(kLaneCostWeight * F(10)), polyline=lane1_p1p2)
    b.state_cost(1, abi.COST_SEMIQUADRATIC_POLYLINE2, dims=pos[1], weight=F(kLaneBoundaryCostWeight * F(10)),
                 polyline=lane1_p1p2, value=-kLaneHalfWidth, flag=0)
    b.state_cost(1, abi.COST_SEMIQUADRATIC_POLYLINE2, dims=pos[1], weight=kLaneBoundaryCostWeight,
                 polyline=lane1_p1p2, value=kLaneHalfWidth, flag=1)
    for pts, oriented_right in pieces:
        b.state_cost(0, abi.COST_SEMIQUADRATIC_POLYLINE2, dims=pos[0], weight=kLaneBoundaryCostWeight,
                     polyline=b.add_polyline(pts), value=0.0, flag=oriented_right)
    for i, w in enumerate((10.0, 1.0)):                                    # :252-266
        b.state_cost(i, abi.COST_QUADRATIC, dims=(offs[i] + 4,), weight=w, value=5.0)
    for i in range(2):                                                     # :269-281
        b.control_cost(i, i, abi.COST_QUADRATIC, dims=(0,), weight=kOmegaCostWeight, value=0.0)
        b.control_cost(i, i, abi.COST_QUADRATIC, dims=(1,), weight=kJerkCostWeight, value=0.0)
    active_from = 10.0 - float(F(0.5))                                     # kTimeHorizon - kFinalTimeWindow (:284-299)
    for i, (gx, gy) in enumerate(((2.5, 50.0), (2.5, -50.0))):
        b.state_cost(i, abi.COST_QUADRATIC, dims=(pos[i][0],), weight=kGoalCostWeight, value=gx, active_from=active_from)
        b.state_cost(i, abi.COST_QUADRATIC, dims=(pos[i][1],), weight=kGoalCostWeight, value=gy, active_from=active_from)
    b.state_cost(0, abi.COST_PROXIMITY, dims=pos[0] + pos[1], weight=kProximityCostWeight, value=kMinProximity)
    b.state_cost(1, abi.COST_PROXIMITY, dims=pos[1] + pos[0], weight=kProximityCostWeight, value=kMinProximity)
    x0 = np.zeros(b.d.xdim, dtype=F)                                       # :164-172
    for i, (x, y, th, v) in enumerate(((2.5, -50.0, math.pi / 2, 10.0), (2.5, 50.0, -math.pi / 2, 2.0))):
        x0[offs[i] + 0], x0[offs[i] + 1], x0[offs[i] + 2], x0[offs[i] + 4] = x, y, F(th), v
    return b.build(), x0


def two_player_collision_params(**overrides) -> abi.SolverParams:
    """SolverParams of exec/two_player_collision/main.cpp:71-74,108-112."""
    base = dict(max_backtracking_steps=100, linesearch=1, expected_decrease_fraction=0.1,
                initial_alpha_scaling=0.75, convergence_tolerance=0.01)
    base.update(overrides)
    return abi.SolverParams.defaults(**base)


def two_player_collision_x0_batch(batch: int, seed: int) -> np.ndarray:
    """Synthetic initial states: the example's, positions moved U(-1, 1) m, headings U(-0.05, 0.05)
    rad, speeds scaled U(0.8, 1.2)."""
    _, x0 = two_player_collision()
    rng = np.random.default_rng(seed)
    out = np.tile(x0, (batch, 1))
    for o in (0, 6):
        out[:, o:o + 2] += rng.uniform(-1.0, 1.0, size=(batch, 2)).astype(F)
        out[:, o + 2] += rng.uniform(-0.05, 0.05, size=batch).astype(F)
        out[:, o + 4] *= rng.uniform(0.8, 1.2, size=batch).astype(F)
    return out.astype(F)


# --------------------------------------------------------------------------
# TwoPlayerCollisionAvoidanceReachabilityExample
# (src/two_player_collision_avoidance_reachability_example.cpp)
# --------------------------------------------------------------------------
def two_player_collision_avoidance_reachability(num_time_steps: int = 100, time_step: float = 0.1,
                                                px0: float = 0.0, py0: float = -5.0):
    """Returns (desc, x0).  2x SinglePlayerCar5D, n = 10, both players max-over-time of a
    SignedDistanceCost between the cars.  CPU oracle only for now (ILQG_DYN_CAR5D and
    ILQG_COST_SIGNED_DISTANCE have no device implementation yet)."""
    b = DescBuilder(num_time_steps, time_step)
    kOmegaCostWeight = 0.1
    heading1, speed = F(0.1), F(5.0)
    for _ in range(2):
        b.add_player(2, 0.0, 0.0, abi.COST_MAX)                            # SetMaxOverTime (:160-161)
    offs = [b.add_subsystem(abi.DYN_CAR5D, 5, i, [4.0]) for i in range(2)]
    # nominal distance (:140-151): where the two cars are at half the horizon if they hold their
    # initial headings and speeds; the double time * float speed narrows to float on the Point2
    half = F(0.5 * (time_step * num_time_steps) * float(speed))
    p1 = (F(F(px0) + F(half * _cosf(heading1))), F(F(py0) + F(half * _sinf(heading1))))
    p2 = (F(F(0.0) + F(half * _cosf(F(0.0)))), F(F(0.0) + F(half * _sinf(F(0.0)))))
    dx, dy = F(p1[0] - p2[0]), F(p1[1] - p2[1])
    nominal = F(np.sqrt(F(F(dx * dx) + F(dy * dy))))
    pos = [(o + 0, o + 1) for o in offs]
    # SignedDistanceCost(dims1, dims2, nominal, "CollisionAvoidance"): the string literal binds to
    # `bool less_is_positive` (a pointer converts to bool before it converts to std::string)
    for i in range(2):
        b.state_cost(i, abi.COST_SIGNED_DISTANCE, dims=pos[0] + pos[1], weight=1.0, value=float(nominal), flag=1)
        b.control_cost(i, i, abi.COST_QUADRATIC, dims=(-1,), weight=kOmegaCostWeight, value=0.0)
    x0 = np.zeros(b.d.xdim, dtype=F)                                       # :106-117
    x0[offs[0] + 0], x0[offs[0] + 1], x0[offs[0] + 2], x0[offs[0] + 4] = px0, py0, heading1, speed
    x0[offs[1] + 4] = speed
    return b.build(), x0


def two_player_collision_avoidance_reachability_params(**overrides) -> abi.SolverParams:
    """SolverParams of exec/two_player_collision_avoidance_reachability_example/main.cpp:76-79,
    114-121 (its two regularization fields are dead)."""
    base = dict(max_backtracking_steps=100, linesearch=1, expected_decrease_fraction=0.1,
                initial_alpha_scaling=0.1, convergence_tolerance=0.01)
    base.update(overrides)
    return abi.SolverParams.defaults(**base)


def two_player_collision_avoidance_reachability_x0_batch(batch: int, seed: int) -> np.ndarray:
    """Synthetic initial states: the example's, positions moved U(-1, 1) m, headings U(-0.1, 0.1)
    rad, speeds scaled U(0.8, 1.2)."""
    _, x0 = two_player_collision_avoidance_reachability()
    rng = np.random.default_rng(seed)
    out = np.tile(x0, (batch, 1))
    for o in (0, 5):
        out[:, o:o + 2] += rng.uniform(-1.0, 1.0, size=(batch, 2)).astype(F)
        out[:, o + 2] += rng.uniform(-0.1, 0.1, size=batch).astype(F)
        out[:, o + 4] *= rng.uniform(0.8, 1.2, size=batch).astype(F)
    return out.astype(F)


# --------------------------------------------------------------------------
# ThreePlayerCollisionAvoidanceReachabilityExample
# (src/three_player_collision_avoidance_reachability_example.cpp)
# --------------------------------------------------------------------------
def three_player_collision_avoidance_reachability(num_time_steps: int = 100, time_step: float = 0.1,
                                                  d0: float = 5.0, v0: float = 5.0, buffer: float = 3.0):
    """Returns (desc, x0).  3x SinglePlayerCar5D heading for the origin, n = 15; each player's cost
    is the max over time of an ExtremeValueCost (max of two SignedDistanceCosts to the other cars)
    plus a control cost, with box constraints on both controls.  CPU oracle only for now."""
    b = DescBuilder(num_time_steps, time_step)
    kOmegaMax, kAMax, kControlCostWeight = 1.0, 0.1, 0.1
    for _ in range(3):
        b.add_player(2, 0.0, 0.0, abi.COST_MAX)                            # SetMaxOverTime (:216-218)
    offs = [b.add_subsystem(abi.DYN_CAR5D, 5, i, [4.0]) for i in range(3)]
    pos = [(o + 0, o + 1) for o in offs]
    # ExtremeValueCost({a, b}, kTakeMin = false) per player (:187-213): one group per player
    members = {0: ((0, 1), (0, 2)), 1: ((0, 1), (1, 2)), 2: ((1, 2), (0, 2))}
    for i in range(3):
        for (a, c) in members[i]:
            b.state_cost(i, abi.COST_SIGNED_DISTANCE, dims=pos[a] + pos[c], weight=1.0, value=buffer, flag=1,
                         group=i + 1, group_is_min=0)
        b.control_cost(i, i, abi.COST_QUADRATIC, dims=(-1,), weight=kControlCostWeight, value=0.0)   # :135-139
        for dim, bound in ((0, kOmegaMax), (1, kAMax)):                                              # :141-185
            b.control_constraint(i, i, abi.CONSTRAINT_SINGLE_DIMENSION, dims=(dim,), value=bound, flag=1)
            b.control_constraint(i, i, abi.CONSTRAINT_SINGLE_DIMENSION, dims=(dim,), value=-bound, flag=0)
    x0 = np.zeros(b.d.xdim, dtype=F)                                       # :106-124, double arithmetic narrowed on store
    kAnglePerturbation = float(F(0.1))
    starts = [(d0, 0.0, -math.pi + kAnglePerturbation),
              (-0.5 * d0, 0.5 * math.sqrt(3.0) * d0, -math.pi / 3.0 + kAnglePerturbation),
              (-0.5 * d0, -0.5 * math.sqrt(3.0) * d0, math.pi / 3.0 + kAnglePerturbation)]
    for i, (x, y, th) in enumerate(starts):
        x0[offs[i] + 0], x0[offs[i] + 1], x0[offs[i] + 2], x0[offs[i] + 4] = x, y, th, v0
    return b.build(), x0


def three_player_collision_avoidance_reachability_params(**overrides) -> abi.SolverParams:
    """SolverParams of exec/three_player_collision_avoidance_reachability_example/main.cpp:76-79,
    114-121 (its two regularization fields are dead)."""
    base = dict(max_backtracking_steps=100, linesearch=1, expected_decrease_fraction=0.1,
                initial_alpha_scaling=0.1, convergence_tolerance=0.01)
    base.update(overrides)
    return abi.SolverParams.defaults(**base)


def three_player_collision_avoidance_reachability_x0_batch(batch: int, seed: int) -> np.ndarray:
    """Synthetic initial states: the example's, positions moved U(-0.5, 0.5) m, headings
    U(-0.1, 0.1) rad, speeds scaled U(0.8, 1.2)."""
    _, x0 = three_player_collision_avoidance_reachability()
    rng = np.random.default_rng(seed)
    out = np.tile(x0, (batch, 1))
    for o in (0, 5, 10):
        out[:, o:o + 2] += rng.uniform(-0.5, 0.5, size=(batch, 2)).astype(F)
        out[:, o + 2] += rng.uniform(-0.1, 0.1, size=batch).astype(F)
        out[:, o + 4] *= rng.uniform(0.8, 1.2, size=batch).astype(F)
    return out.astype(F)


# --------------------------------------------------------------------------
# OnePlayerReachabilityExample (src/one_player_reachability_example.cpp)
# --------------------------------------------------------------------------
def one_player_reachability(num_time_steps: int = 100, time_step: float = 0.1, px0: float = 1.75,
                            py0: float = 1.75, theta0: float = 0.0):
    """Returns (desc, x0).  ONE player: a SinglePlayerDubinsCar (n = 3, m = 1) avoiding a disc,
    max-over-time signed distance to its 10-gon boundary, turn rate boxed by two constraints.
    CPU oracle only for now (ILQG_DYN_DUBINS has no device implementation yet)."""
    b = DescBuilder(num_time_steps, time_step)
    kTargetRadius, kOmegaMax, kOmegaCostWeight, kSpeed = 2.0, 1.0, 0.1, 1.0
    b.add_player(1, 0.0, 0.0, abi.COST_MAX)                                # SetMaxOverTime
    b.add_subsystem(abi.DYN_DUBINS, 3, 0, [kSpeed])
    circle = b.add_polyline(draw_circle((0.0, 0.0), kTargetRadius, 10))
    # Polyline2SignedDistanceCost(circle, dims, kAvoid, "Target"): as in Air3DExample the bool lands
    # in `float nominal` (1.0) and the string literal in `bool oriented_same_as_polyline` (true)
    b.state_cost(0, abi.COST_POLYLINE2_SIGNED_DISTANCE, dims=(0, 1), polyline=circle, value=1.0, flag=1)
    b.control_cost(0, 0, abi.COST_QUADRATIC, dims=(-1,), weight=kOmegaCostWeight, value=0.0)
    b.control_constraint(0, 0, abi.CONSTRAINT_SINGLE_DIMENSION, dims=(0,), value=kOmegaMax, flag=1)
    b.control_constraint(0, 0, abi.CONSTRAINT_SINGLE_DIMENSION, dims=(0,), value=-kOmegaMax, flag=0)
    x0 = np.array([px0, py0, theta0], dtype=F)
    return b.build(), x0


def one_player_reachability_params(**overrides) -> abi.SolverParams:
    """SolverParams of exec/one_player_reachability_example/main.cpp (same flags as the other
    reachability executables)."""
    base = dict(max_backtracking_steps=100, linesearch=1, expected_decrease_fraction=0.1,
                initial_alpha_scaling=0.1, convergence_tolerance=0.01)
    base.update(overrides)
    return abi.SolverParams.defaults(**base)


def one_player_reachability_x0_batch(batch: int, seed: int) -> np.ndarray:
    """Synthetic initial states: positions U(-4, 4)^2 outside the disc's radius + 0.5, headings U(-pi, pi)."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < batch:
        x, y = rng.uniform(-4.0, 4.0, size=2)
        if math.hypot(x, y) > 2.5:
            out.append((x, y, rng.uniform(-math.pi, math.pi)))
    return np.array(out, dtype=F)


# --------------------------------------------------------------------------
# DubinsOriginExample (src/dubins_origin_example.cpp)
# --------------------------------------------------------------------------
def dubins_origin(num_time_steps: int = 100, time_step: float = 0.1):
    """Returns (desc, x0).  2x SinglePlayerDubinsCar, n = 6, m = (1, 1): P1 wants P2 at the origin,
    P2 wants to be where P1 is (QuadraticDifferenceCost).  Its executable runs WITHOUT the
    linesearch (SolverParams::linesearch = false).  CPU oracle only for now."""
    b = DescBuilder(num_time_steps, time_step)
    kOmegaCostWeight, kAttractionCostWeight, kGoalCostWeight, kSpeed = 100.0, 10.0, 10.0, 1.0
    for _ in range(2):
        b.add_player(1, 0.0, 0.0)
    offs = [b.add_subsystem(abi.DYN_DUBINS, 3, i, [kSpeed]) for i in range(2)]
    p1, p2 = (offs[0] + 0, offs[0] + 1), (offs[1] + 0, offs[1] + 1)
    b.state_cost(1, abi.COST_QUADRATIC_DIFFERENCE, dims=p1 + p2, weight=kAttractionCostWeight, flag=2)   # :124-129
    b.control_cost(0, 0, abi.COST_QUADRATIC, dims=(0,), weight=kOmegaCostWeight, value=0.0)               # :132-138
    b.control_cost(1, 1, abi.COST_QUADRATIC, dims=(0,), weight=kOmegaCostWeight, value=0.0)
    b.state_cost(0, abi.COST_QUADRATIC, dims=(p2[0],), weight=kGoalCostWeight, value=0.0)                 # :141-146
    b.state_cost(0, abi.COST_QUADRATIC, dims=(p2[1],), weight=kGoalCostWeight, value=0.0)
    x0 = np.zeros(b.d.xdim, dtype=F)                                       # :70-75, :107-114
    x0[offs[0] + 1], x0[offs[0] + 2] = -10.0, F(math.pi - 0.01)
    x0[offs[1] + 1], x0[offs[1] + 2] = 10.0, F(1.5 * math.pi)
    return b.build(), x0


def dubins_origin_params(**overrides) -> abi.SolverParams:
    """SolverParams of exec/dubins_origin_example/main.cpp:72-75,110-114: no linesearch."""
    base = dict(max_backtracking_steps=100, linesearch=0, expected_decrease_fraction=0.1,
                initial_alpha_scaling=0.1, convergence_tolerance=0.1)
    base.update(overrides)
    return abi.SolverParams.defaults(**base)


def dubins_origin_x0_batch(batch: int, seed: int) -> np.ndarray:
    """Synthetic initial states: the example's, positions moved U(-2, 2) m, headings U(-0.3, 0.3) rad."""
    _, x0 = dubins_origin()
    rng = np.random.default_rng(seed)
    out = np.tile(x0, (batch, 1))
    for o in (0, 3):
        out[:, o:o + 2] += rng.uniform(-2.0, 2.0, size=(batch, 2)).astype(F)
        out[:, o + 2] += rng.uniform(-0.3, 0.3, size=batch).astype(F)
    return out.astype(F)


# --------------------------------------------------------------------------
# TwoPlayerReachabilityExample (src/two_player_reachability_example.cpp)
# --------------------------------------------------------------------------
def two_player_reachability(num_time_steps: int = 100, time_step: float = 0.1, px0: float = 0.0,
                            py0: float = -10.0, theta0: float = math.pi / 4.0, v0: float = 5.0):
    """Returns (desc, x0).  TwoPlayerUnicycle4D (one coupled subsystem, n = 4, m = (2, 2)): P1
    maximises over time, P2 minimises over time the signed distance to a disc.  Unconstrained.
    CPU oracle only for now."""
    b = DescBuilder(num_time_steps, time_step)
    kControlCostWeight, kTargetRadius = 0.1, 1.0
    b.add_player(2, 0.0, 0.0, abi.COST_MAX)                                # p1_cost.SetMaxOverTime()
    b.add_player(2, 0.0, 0.0, abi.COST_MIN)                                # p2_cost.SetMinOverTime()
    b.add_subsystem(abi.DYN_TWO_PLAYER_UNICYCLE4D, 4, 0, [])
    circle = b.add_polyline(draw_circle((0.0, 0.0), kTargetRadius, 10))
    # Polyline2SignedDistanceCost(circle, dims, !kReach / kReach, "Target"): bool -> float nominal,
    # string literal -> bool oriented_same_as_polyline, as in Air3DExample
    b.state_cost(0, abi.COST_POLYLINE2_SIGNED_DISTANCE, dims=(0, 1), polyline=circle, value=0.0, flag=1)
    b.state_cost(1, abi.COST_POLYLINE2_SIGNED_DISTANCE, dims=(0, 1), polyline=circle, value=1.0, flag=1)
    b.control_cost(0, 0, abi.COST_QUADRATIC, dims=(-1,), weight=kControlCostWeight, value=0.0)
    b.control_cost(1, 1, abi.COST_QUADRATIC, dims=(-1,), weight=kControlCostWeight, value=0.0)
    x0 = np.array([px0, py0, theta0, v0], dtype=F)
    return b.build(), x0


def two_player_reachability_params(**overrides) -> abi.SolverParams:
    """SolverParams of exec/two_player_reachability_example/main.cpp:75-78."""
    base = dict(max_backtracking_steps=100, linesearch=1, expected_decrease_fraction=0.1,
                initial_alpha_scaling=0.1, convergence_tolerance=0.01)
    base.update(overrides)
    return abi.SolverParams.defaults(**base)


def two_player_reachability_x0_batch(batch: int, seed: int) -> np.ndarray:
    """Synthetic initial states: positions U(-10, 10)^2 outside radius 2, headings U(-pi, pi), speeds U(3, 6)."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < batch:
        x, y = rng.uniform(-10.0, 10.0, size=2)
        if math.hypot(x, y) > 2.0:
            out.append((x, y, rng.uniform(-math.pi, math.pi), rng.uniform(3.0, 6.0)))
    return np.array(out, dtype=F)


# --------------------------------------------------------------------------
# ModifiedAir3DExample (src/modified_air_3d_example.cpp)
# --------------------------------------------------------------------------
def modified_air_3d(num_time_steps: int = 100, time_step: float = 0.1, rx0: float = 4.0, ry0: float = 3.0,
                    rtheta0: float = math.pi / 4.0, ve: float = 1.0, vp: float = 1.0):
    """Returns (desc, x0).  2x SinglePlayerPointMass2D (n = 8, m = (2, 2)), pursuit-evasion written
    with QuadraticDifferenceCosts of weight -1e6 / +1e6; state regularization 1.  CPU oracle only
    for now."""
    b = DescBuilder(num_time_steps, time_step)
    kOmegaCostWeight = 0.1
    for _ in range(2):
        b.add_player(2, 1.0, 0.0)                                          # PlayerCost("P1", 1.0, 0.0)
    offs = [b.add_subsystem(abi.DYN_POINT_MASS_2D, 4, i, []) for i in range(2)]
    p1, p2 = (offs[0] + 0, offs[0] + 1), (offs[1] + 0, offs[1] + 1)
    for i, w in enumerate((-1e6, 1e6)):                                    # kEvaderWeight, kPursuerWeight
        b.state_cost(i, abi.COST_QUADRATIC_DIFFERENCE, dims=p1 + p2, weight=w, flag=2)
        b.control_cost(i, i, abi.COST_QUADRATIC, dims=(-1,), weight=kOmegaCostWeight, value=0.0)
    x0 = np.zeros(b.d.xdim, dtype=F)                                       # double arithmetic, narrowed on store
    x0[offs[0] + 2] = ve
    x0[offs[1] + 0], x0[offs[1] + 1] = rx0, ry0
    x0[offs[1] + 2], x0[offs[1] + 3] = vp * math.cos(rtheta0), vp * math.sin(rtheta0)
    return b.build(), x0


def modified_air_3d_params(**overrides) -> abi.SolverParams:
    """SolverParams of exec/modified_air_3d_example/main.cpp:75-78."""
    base = dict(max_backtracking_steps=100, linesearch=1, expected_decrease_fraction=0.1,
                initial_alpha_scaling=0.1, convergence_tolerance=0.01)
    base.update(overrides)
    return abi.SolverParams.defaults(**base)


def modified_air_3d_x0_batch(batch: int, seed: int) -> np.ndarray:
    """Synthetic initial states: pursuer position U(-6, 6)^2, heading U(-pi, pi), both speeds 1."""
    rng = np.random.default_rng(seed)
    out = np.zeros((batch, 8), dtype=F)
    out[:, 2] = 1.0
    out[:, 4:6] = rng.uniform(-6.0, 6.0, size=(batch, 2))
    th = rng.uniform(-math.pi, math.pi, size=batch)
    out[:, 6], out[:, 7] = np.cos(th), np.sin(th)
    return out


# --------------------------------------------------------------------------
# Air3DExample
# --------------------------------------------------------------------------
def draw_circle(center, radius, num_segments):
    """DrawCircle, src/draw_shapes.cpp:62-73."""
    pts = [(center[0] + radius, center[1] + 0.0)]
    for ii in range(num_segments):
        angle = F(2.0 * math.pi * float(F(ii + 1)) / float(F(num_segments)))
        pts.append((float(F(center[0]) + F(radius) * _cosf(angle)),
                    float(F(center[1]) + F(radius) * _sinf(angle))))
    return pts


def air_3d(num_time_steps: int = 100, time_step: float = 0.1, rx0=4.0, ry0=3.0,
           rtheta0=math.pi / 4.0, ve=1.0, vp=1.0):
    """Returns (desc, x0). Two-player zero-sum-like pursuit-evasion, n = 3, m = (1,1)."""
    b = DescBuilder(num_time_steps, time_step)
    kOmegaCostWeight, kOmegaMax, kTargetRadius = 0.1, 1.0, 5.0
    b.add_player(1, 0.0, 0.0, abi.COST_MAX)  # p1_cost.SetMaxOverTime() :139
    b.add_player(1, 0.0, 0.0, abi.COST_MIN)  # p2_cost.SetMinOverTime() :140
    b.add_subsystem(abi.DYN_AIR3D, 3, 0, [ve, vp])
    circle = b.add_polyline(draw_circle((0.0, 0.0), kTargetRadius, 10))
    # Polyline2SignedDistanceCost(circle, {rx,ry}, !kReach, "Target"): the bool lands in
    # `float nominal` and the string literal in `bool oriented_same_as_polyline`:
    # nominal = 0.0 (P1) / 1.0 (P2), oriented = true for both.
    b.state_cost(0, abi.COST_POLYLINE2_SIGNED_DISTANCE, dims=(0, 1), polyline=circle, value=0.0,
                 flag=1)
    b.state_cost(1, abi.COST_POLYLINE2_SIGNED_DISTANCE, dims=(0, 1), polyline=circle, value=1.0,
                 flag=1)
    b.control_cost(0, 0, abi.COST_QUADRATIC, dims=(-1,), weight=kOmegaCostWeight, value=0.0)
    b.control_cost(1, 1, abi.COST_QUADRATIC, dims=(-1,), weight=kOmegaCostWeight, value=0.0)
    b.control_constraint(0, 0, abi.CONSTRAINT_SINGLE_DIMENSION, dims=(0,), value=kOmegaMax, flag=1)
    b.control_constraint(0, 0, abi.CONSTRAINT_SINGLE_DIMENSION, dims=(0,), value=-kOmegaMax, flag=0)
    b.control_constraint(1, 1, abi.CONSTRAINT_SINGLE_DIMENSION, dims=(0,), value=kOmegaMax, flag=1)
    b.control_constraint(1, 1, abi.CONSTRAINT_SINGLE_DIMENSION, dims=(0,), value=-kOmegaMax, flag=0)
    x0 = np.array([rx0, ry0, rtheta0], dtype=F)
    return b.build(), x0


def air_3d_params(**overrides) -> abi.SolverParams:
    """SolverParams of exec/air_3d_example/main.cpp:75-78,111-118 (its two
    regularization fields are dead)."""
    base = dict(max_backtracking_steps=100, linesearch=1, expected_decrease_fraction=0.1,
                initial_alpha_scaling=0.1, convergence_tolerance=0.01)
    base.update(overrides)
    return abi.SolverParams.defaults(**base)


def air_3d_x0_grid(side: int = 128) -> np.ndarray:
    """SURVEY section 8d input #4: side x side grid r_x,r_y in [-6,6], r_theta = pi/4."""
    g = np.linspace(-6.0, 6.0, side, dtype=F)
    rx, ry = np.meshgrid(g, g, indexing="ij")
    out = np.stack([rx.ravel(), ry.ravel(), np.full(side * side, math.pi / 4.0, dtype=F)], axis=1)
    r = np.hypot(out[:, 0], out[:, 1])
    out[r < 0.5, 0] += 1.0  # exclude the origin neighbourhood
    return out.astype(F)


# --------------------------------------------------------------------------
# LQ-only handle: the system of test/test_lq_solver.cpp
# --------------------------------------------------------------------------
def lq_only(num_time_steps: int, xdim: int, udims: Sequence[int],
            cross_pairs: Sequence[Tuple[int, int]] = ()):
    """Descriptor with no dynamics/costs: lin/quad are supplied with upload_lq.
    cross_pairs lists extra (i, j) control blocks R_ij, i != j."""
    b = DescBuilder(num_time_steps, 0.1)
    for m in udims:
        b.add_player(m)
    b.d.xdim = xdim
    for (i, j) in cross_pairs:
        # a zero-weight control cost only declares that the (i,j) block exists
        b.control_cost(i, j, abi.COST_QUADRATIC, dims=(-1,), weight=0.0, value=0.0)
    return b.build()
