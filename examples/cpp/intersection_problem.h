// The three-player intersection game (two cars and a pedestrian-like unicycle crossing each
// other's lanes) written against the ilqgames class API as hosted by this repository.  Same game
// as the reference's ThreePlayerIntersectionExample (src/three_player_intersection_example.cpp:
// dynamics :164-169, initial state :171-185, costs :187-394) and as
// ilqgames_b200/problems.py:three_player_intersection -- the numbers are that game's definition.
#ifndef ILQGAMES_B200_EXAMPLES_INTERSECTION_PROBLEM_H
#define ILQGAMES_B200_EXAMPLES_INTERSECTION_PROBLEM_H

#include <ilqgames/cost/quadratic_cost.h>
#include <ilqgames/cost/quadratic_polyline2_cost.h>
#include <ilqgames/constraint/proximity_constraint.h>
#include <ilqgames/dynamics/concatenated_dynamical_system.h>
#include <ilqgames/dynamics/single_player_car_6d.h>
#include <ilqgames/dynamics/single_player_unicycle_4d.h>
#include <ilqgames/geometry/polyline2.h>
#include <ilqgames/solver/top_down_renderable_problem.h>

#include <cmath>

namespace ilqgames_b200_examples {

using namespace ilqgames;

class IntersectionProblem : public TopDownRenderableProblem {
 public:
  IntersectionProblem() : TopDownRenderableProblem() {}

  struct Agent {
    bool is_car;          // car6d (x, y, theta, phi, v, a) or unicycle4d (x, y, theta, v)
    float x, y, heading, speed, nominal_speed;
    PointList2 lane;
  };

  static std::vector<Agent> Agents() {
    const float x1 = -2.0f, x2 = -10.0f, y3 = 16.0f;
    return {
        {true, x1, -30.0f, (float)M_PI_2, 4.0f, 8.0f, {Point2(x1, -1000.0f), Point2(x1, 1000.0f)}},
        {true, x2, 45.0f, (float)-M_PI_2, 3.0f, 5.0f,
         {Point2(x2, 1000.0f), Point2(x2, 18.0f), Point2(x2 + 0.5f, 15.0f), Point2(x2 + 1.0f, 14.0f),
          Point2(x2 + 3.0f, 12.5f), Point2(x2 + 6.0f, 12.0f), Point2(1000.0f, 12.0f)}},
        {false, -11.0f, y3, 0.0f, 1.25f, 1.5f, {Point2(-1000.0f, y3), Point2(1000.0f, y3)}},
    };
  }

  void ConstructDynamics() override {
    SubsystemList subsystems;
    for (const Agent& a : Agents()) {
      if (a.is_car)
        subsystems.push_back(std::make_shared<SinglePlayerCar6D>(4.0f /* inter-axle distance, m */));
      else
        subsystems.push_back(std::make_shared<SinglePlayerUnicycle4D>());
    }
    dynamics_.reset(new ConcatenatedDynamicalSystem(subsystems));
  }

  void ConstructInitialState() override {
    x0_ = VectorXf::Zero(dynamics_->XDim());
    const auto agents = Agents();
    for (size_t i = 0; i < agents.size(); i++) {
      const Dimension s = Start(i);
      x0_(s + 0) = agents[i].x;
      x0_(s + 1) = agents[i].y;
      x0_(s + 2) = agents[i].heading;
      x0_(s + SpeedIdx(agents[i])) = agents[i].speed;
    }
  }

  void ConstructPlayerCosts() override {
    const auto agents = Agents();
    for (size_t i = 0; i < agents.size(); i++)
      player_costs_.emplace_back("P" + std::to_string(i + 1), 1.0f /* state reg */, 5.0f /* control reg */);
    for (size_t i = 0; i < agents.size(); i++) {
      PlayerCost& pc = player_costs_[i];
      const Dimension s = Start(i);
      const std::pair<Dimension, Dimension> xy(s, s + 1);
      pc.AddStateCost(std::make_shared<QuadraticPolyline2Cost>(25.0f, Polyline2(agents[i].lane), xy, "LaneCenter"));
      pc.AddStateCost(std::make_shared<QuadraticCost>(100.0f, s + SpeedIdx(agents[i]), agents[i].nominal_speed, "NominalV"));
      if (WithLaneBoundaries()) {
        // the lane as a hard constraint: stay within 2.5 m of its centre line on either side (the
        // reference example constructs these and leaves them commented out,
        // src/three_player_intersection_example.cpp:214-251)
        pc.AddStateConstraint(std::make_shared<Polyline2SignedDistanceConstraint>(Polyline2(agents[i].lane), xy, 2.5f, false, "LaneRightBoundary"));
        pc.AddStateConstraint(std::make_shared<Polyline2SignedDistanceConstraint>(Polyline2(agents[i].lane), xy, -2.5f, true, "LaneLeftBoundary"));
      }
      // controls: (steering or turn rate, jerk or acceleration), both lightly penalised
      pc.AddControlCost((PlayerIndex)i, std::make_shared<QuadraticCost>(0.1f, 0, 0.0f, "Steering"));
      pc.AddControlCost((PlayerIndex)i, std::make_shared<QuadraticCost>(0.1f, 1, 0.0f, agents[i].is_car ? "Jerk" : "Acceleration"));
      // keep at least 6 m from each of the other two
      for (size_t j = 0; j < agents.size() && WithProximityConstraints(); j++) {
        if (j == i) continue;
        const std::pair<Dimension, Dimension> other(Start(j), Start(j) + 1);
        pc.AddStateConstraint(std::make_shared<ProximityConstraint>(xy, other, 6.0f, false, "ProximityConstraintP" + std::to_string(j + 1)));
      }
    }
  }

  // false: the same game without the proximity constraints (used by the receding-horizon test)
  virtual bool WithProximityConstraints() const { return true; }
  // true: each player's lane boundaries as Polyline2SignedDistanceConstraints
  virtual bool WithLaneBoundaries() const { return false; }

  std::vector<float> Xs(const VectorXf& x) const override { return Pick(x, 0); }
  std::vector<float> Ys(const VectorXf& x) const override { return Pick(x, 1); }
  std::vector<float> Thetas(const VectorXf& x) const override { return Pick(x, 2); }

 private:
  static Dimension SpeedIdx(const Agent& a) { return a.is_car ? SinglePlayerCar6D::kVIdx : SinglePlayerUnicycle4D::kVIdx; }
  Dimension Start(size_t i) const {
    return static_cast<const ConcatenatedDynamicalSystem*>(dynamics_.get())->SubsystemStartDim((PlayerIndex)i);
  }
  std::vector<float> Pick(const VectorXf& x, Dimension offset) const {
    std::vector<float> out;
    for (size_t i = 0; i < Agents().size(); i++) out.push_back(x(Start(i) + offset));
    return out;
  }
};

}  // namespace ilqgames_b200_examples

#endif
