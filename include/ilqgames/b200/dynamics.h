// dynamics.h -- the in-scope dynamical systems as descriptors (kind + parameters).  Integration
// and linearization run on the device (csrc/ilqg_device.cuh: subsystem_xdot,
// subsystem_integrate, subsystem_linearize_sink); these classes keep the reference's names,
// constructor arguments and index constants so problem definitions compile unchanged.
#ifndef ILQGAMES_B200_DYNAMICS_H
#define ILQGAMES_B200_DYNAMICS_H

#include <ilqg.h>
#include <ilqgames/b200/core.h>

namespace ilqgames {

// include/ilqgames/dynamics/single_player_dynamical_system.h:55-92
class SinglePlayerDynamicalSystem {
 public:
  virtual ~SinglePlayerDynamicalSystem() {}
  Dimension XDim() const { return xdim_; }
  Dimension UDim() const { return udim_; }
  virtual std::vector<Dimension> PositionDimensions() const = 0;
  // device kind + parameters (ILQG_DYN_*); false = no device implementation
  virtual bool Describe(ilqg_subsystem_desc* /*out*/) const { return false; }

 protected:
  SinglePlayerDynamicalSystem(Dimension xdim, Dimension udim) : xdim_(xdim), udim_(udim) {}
  const Dimension xdim_;
  const Dimension udim_;
};

// include/ilqgames/dynamics/single_player_car_6d.h:62-138; state (x, y, theta, phi, v, a),
// controls (steering rate, jerk)
class SinglePlayerCar6D : public SinglePlayerDynamicalSystem {
 public:
  SinglePlayerCar6D(float inter_axle_distance)
      : SinglePlayerDynamicalSystem(kNumXDims, kNumUDims), inter_axle_distance_(inter_axle_distance) {}
  std::vector<Dimension> PositionDimensions() const override { return {kPxIdx, kPyIdx}; }
  bool Describe(ilqg_subsystem_desc* out) const override {
    out->kind = ILQG_DYN_CAR6D;
    out->params[0] = inter_axle_distance_;
    return true;
  }
  static constexpr Dimension kNumXDims = 6, kPxIdx = 0, kPyIdx = 1, kThetaIdx = 2, kPhiIdx = 3, kVIdx = 4, kAIdx = 5;
  static constexpr Dimension kNumUDims = 2, kOmegaIdx = 0, kJerkIdx = 1;

 private:
  const float inter_axle_distance_;
};

// include/ilqgames/dynamics/single_player_unicycle_4d.h:60-116; state (x, y, theta, v)
class SinglePlayerUnicycle4D : public SinglePlayerDynamicalSystem {
 public:
  SinglePlayerUnicycle4D() : SinglePlayerDynamicalSystem(kNumXDims, kNumUDims) {}
  std::vector<Dimension> PositionDimensions() const override { return {kPxIdx, kPyIdx}; }
  bool Describe(ilqg_subsystem_desc* out) const override {
    out->kind = ILQG_DYN_UNICYCLE4D;
    return true;
  }
  static constexpr Dimension kNumXDims = 4, kPxIdx = 0, kPyIdx = 1, kThetaIdx = 2, kVIdx = 3;
  static constexpr Dimension kNumUDims = 2, kOmegaIdx = 0, kAIdx = 1;
};

// include/ilqgames/dynamics/single_player_car_5d.h:60-147; state (x, y, theta, phi, v), controls
// (steering rate, acceleration).  Described as ILQG_DYN_CAR5D, which only the CPU oracle
// implements so far (the CUDA library refuses it).
class SinglePlayerCar5D : public SinglePlayerDynamicalSystem {
 public:
  SinglePlayerCar5D(float inter_axle_distance)
      : SinglePlayerDynamicalSystem(kNumXDims, kNumUDims), inter_axle_distance_(inter_axle_distance) {}
  std::vector<Dimension> PositionDimensions() const override { return {kPxIdx, kPyIdx}; }
  bool Describe(ilqg_subsystem_desc* out) const override {
    out->kind = ILQG_DYN_CAR5D;
    out->params[0] = inter_axle_distance_;
    return true;
  }
  static constexpr Dimension kNumXDims = 5, kPxIdx = 0, kPyIdx = 1, kThetaIdx = 2, kPhiIdx = 3, kVIdx = 4;
  static constexpr Dimension kNumUDims = 2, kOmegaIdx = 0, kAIdx = 1;

 private:
  const float inter_axle_distance_;
};

// include/ilqgames/dynamics/single_player_dubins_car.h:56-118; state (x, y, theta) at constant
// speed, control = turn rate.  ILQG_DYN_DUBINS: CPU oracle only so far.
class SinglePlayerDubinsCar : public SinglePlayerDynamicalSystem {
 public:
  SinglePlayerDubinsCar(float v) : SinglePlayerDynamicalSystem(kNumXDims, kNumUDims), v_(v) { CHECK_GT(v_, 0.0); }
  std::vector<Dimension> PositionDimensions() const override { return {kPxIdx, kPyIdx}; }
  bool Describe(ilqg_subsystem_desc* out) const override {
    out->kind = ILQG_DYN_DUBINS;
    out->params[0] = v_;
    return true;
  }
  static constexpr Dimension kNumXDims = 3, kPxIdx = 0, kPyIdx = 1, kThetaIdx = 2;
  static constexpr Dimension kNumUDims = 1, kOmegaIdx = 0;

 private:
  const float v_;
};

// include/ilqgames/dynamics/single_player_point_mass_2d.h:56-118; state (x, y, vx, vy), controls
// (ax, ay).  ILQG_DYN_POINT_MASS_2D: CPU oracle only so far.
class SinglePlayerPointMass2D : public SinglePlayerDynamicalSystem {
 public:
  SinglePlayerPointMass2D() : SinglePlayerDynamicalSystem(kNumXDims, kNumUDims) {}
  std::vector<Dimension> PositionDimensions() const override { return {kPxIdx, kPyIdx}; }
  bool Describe(ilqg_subsystem_desc* out) const override {
    out->kind = ILQG_DYN_POINT_MASS_2D;
    return true;
  }
  static constexpr Dimension kNumXDims = 4, kPxIdx = 0, kPyIdx = 1, kVxIdx = 2, kVyIdx = 3;
  static constexpr Dimension kNumUDims = 2, kAxIdx = 0, kAyIdx = 1;
};

// include/ilqgames/dynamics/multi_player_integrable_system.h:58-140
class MultiPlayerIntegrableSystem {
 public:
  virtual ~MultiPlayerIntegrableSystem() {}
  virtual bool TreatAsLinear() const { return false; }
  Dimension XDim() const { return xdim_; }
  Dimension TotalUDim() const {
    Dimension total = 0;
    for (PlayerIndex ii = 0; ii < NumPlayers(); ii++) total += UDim(ii);
    return total;
  }
  virtual Dimension UDim(PlayerIndex player_idx) const = 0;
  virtual PlayerIndex NumPlayers() const = 0;
  virtual std::vector<Dimension> PositionDimensions() const = 0;
  // subsystem table of ilqg_problem_desc; false = not describable
  virtual bool Describe(ilqg_problem_desc* /*desc*/) const { return false; }

  // src/multi_player_integrable_system.cpp:54-83: the state x0, taken at time t0, carried to time t
  // under the given plan (u_i = u_ref - P_i (x - x_ref) - alpha_i).  Runs on the device
  // (ilqg_integrate_plan); defined in solvers.h below b200::Handle.
  VectorXf Integrate(Time t0, Time t, const VectorXf& x0, const OperatingPoint& operating_point,
                     const std::vector<Strategy>& strategies) const;

 protected:
  MultiPlayerIntegrableSystem(Dimension xdim) : xdim_(xdim) {}
  const Dimension xdim_;
};

// include/ilqgames/dynamics/multi_player_dynamical_system.h
class MultiPlayerDynamicalSystem : public MultiPlayerIntegrableSystem {
 protected:
  MultiPlayerDynamicalSystem(Dimension xdim) : MultiPlayerIntegrableSystem(xdim) {}
};

// include/ilqgames/dynamics/concatenated_dynamical_system.h:56-120
using SubsystemList = PtrVector<SinglePlayerDynamicalSystem>;

class ConcatenatedDynamicalSystem : public MultiPlayerDynamicalSystem {
 public:
  ConcatenatedDynamicalSystem(const SubsystemList& subsystems)
      : MultiPlayerDynamicalSystem(SumXDims(subsystems)), subsystems_(subsystems) {
    Dimension start = 0;
    for (const auto& s : subsystems_) {
      subsystem_start_dims_.push_back(start);
      start += s->XDim();
    }
  }
  const SubsystemList& Subsystems() const { return subsystems_; }
  PlayerIndex NumPlayers() const override { return (PlayerIndex)subsystems_.size(); }
  Dimension SubsystemStartDim(PlayerIndex player_idx) const { return subsystem_start_dims_[player_idx]; }
  Dimension SubsystemXDim(PlayerIndex player_idx) const { return subsystems_[player_idx]->XDim(); }
  Dimension UDim(PlayerIndex player_idx) const override { return subsystems_[player_idx]->UDim(); }
  std::vector<Dimension> PositionDimensions() const override {
    std::vector<Dimension> dims;
    for (size_t ii = 0; ii < subsystems_.size(); ii++)
      for (Dimension d : subsystems_[ii]->PositionDimensions()) dims.push_back(subsystem_start_dims_[ii] + d);
    return dims;
  }
  bool Describe(ilqg_problem_desc* desc) const override {
    if (subsystems_.size() > ILQG_MAX_SUBSYSTEMS) return false;
    desc->num_subsystems = (int32_t)subsystems_.size();
    for (size_t ii = 0; ii < subsystems_.size(); ii++) {
      ilqg_subsystem_desc& s = desc->subsystems[ii];
      s = ilqg_subsystem_desc();
      if (!subsystems_[ii]->Describe(&s)) return false;
      s.x_offset = subsystem_start_dims_[ii];
      s.first_player = (int32_t)ii;
    }
    return true;
  }

 private:
  static Dimension SumXDims(const SubsystemList& subsystems) {
    Dimension total = 0;
    for (const auto& s : subsystems) total += CHECK_NOTNULL_PTR(s)->XDim();
    return total;
  }
  template <typename P>
  static const P& CHECK_NOTNULL_PTR(const P& p) { CHECK(p.get() != nullptr); return p; }

  const SubsystemList subsystems_;
  std::vector<Dimension> subsystem_start_dims_;
};

// include/ilqgames/dynamics/air_3d.h:62-149: relative (x, y, heading); player 1 = evader turn
// rate, player 2 = pursuer turn rate; one coupled 3-state subsystem
class Air3D : public MultiPlayerDynamicalSystem {
 public:
  Air3D(float evader_speed, float pursuer_speed)
      : MultiPlayerDynamicalSystem(kNumXDims), evader_speed_(evader_speed), pursuer_speed_(pursuer_speed) {}
  Dimension UDim(PlayerIndex player_idx) const override { return player_idx == 0 ? kNumU1Dims : kNumU2Dims; }
  PlayerIndex NumPlayers() const override { return 2; }
  std::vector<Dimension> PositionDimensions() const override { return {kRxIdx, kRyIdx}; }
  bool Describe(ilqg_problem_desc* desc) const override {
    desc->num_subsystems = 1;
    ilqg_subsystem_desc& s = desc->subsystems[0];
    s = ilqg_subsystem_desc();
    s.kind = ILQG_DYN_AIR3D;
    s.params[0] = evader_speed_;
    s.params[1] = pursuer_speed_;
    return true;
  }
  static constexpr Dimension kNumXDims = 3, kRxIdx = 0, kRyIdx = 1, kRThetaIdx = 2;
  static constexpr Dimension kNumU1Dims = 1, kOmega1Idx = 0, kNumU2Dims = 1, kOmega2Idx = 0;

 private:
  const float evader_speed_, pursuer_speed_;
};

// include/ilqgames/dynamics/two_player_unicycle_4d.h:60-137: one unicycle (x, y, theta, v); player 1
// steers it (omega, a), player 2 pushes its position (dx, dy).  ILQG_DYN_TWO_PLAYER_UNICYCLE4D: CPU
// oracle only so far.
class TwoPlayerUnicycle4D : public MultiPlayerDynamicalSystem {
 public:
  TwoPlayerUnicycle4D() : MultiPlayerDynamicalSystem(kNumXDims) {}
  Dimension UDim(PlayerIndex player_idx) const override { return player_idx == 0 ? kNumU1Dims : kNumU2Dims; }
  PlayerIndex NumPlayers() const override { return kNumPlayers; }
  std::vector<Dimension> PositionDimensions() const override { return {kPxIdx, kPyIdx}; }
  bool Describe(ilqg_problem_desc* desc) const override {
    desc->num_subsystems = 1;
    ilqg_subsystem_desc& s = desc->subsystems[0];
    s = ilqg_subsystem_desc();
    s.kind = ILQG_DYN_TWO_PLAYER_UNICYCLE4D;
    return true;
  }
  static constexpr Dimension kNumXDims = 4, kPxIdx = 0, kPyIdx = 1, kThetaIdx = 2, kVIdx = 3;
  static constexpr PlayerIndex kNumPlayers = 2;
  static constexpr Dimension kNumU1Dims = 2, kOmegaIdx = 0, kAIdx = 1, kNumU2Dims = 2, kDxIdx = 0, kDyIdx = 1;
};

}  // namespace ilqgames

#endif
