// costs.h -- Polyline2, the in-scope Cost / Constraint classes and PlayerCost, re-authored as
// parameter carriers.  The reference evaluates and differentiates these on the CPU through
// virtual calls (src/*_cost.cpp); here that arithmetic lives in the CUDA kernels, and a class
// only has to (1) accept the reference's constructor arguments and (2) describe itself as one
// POD record of include/ilqg.h (`Describe`).  A class without a device implementation returns
// false from Describe, which makes ilqg::DescribeProblem fail loudly (no CPU fallback).
#ifndef ILQGAMES_B200_COSTS_H
#define ILQGAMES_B200_COSTS_H

#include <ilqg.h>
#include <ilqgames/b200/core.h>

#include <cstring>

namespace ilqgames {

// ---- include/ilqgames/geometry/line_segment2.h:52-96: the accessors host code uses to place
// agents on a lane (src/roundabout_merging_example.cpp:195-212); closest-point queries run on the
// device (csrc/ilqg_device.cuh: segment_closest)
class LineSegment2 {
 public:
  LineSegment2(const Point2& point1, const Point2& point2) : p1_(point1), p2_(point2) {
    length_ = (point1 - point2).norm();
    CHECK_GT(length_, constants::kSmallNumber);
    unit_direction_ = (point2 - point1) / length_;
  }
  float Length() const { return length_; }
  const Point2& FirstPoint() const { return p1_; }
  const Point2& SecondPoint() const { return p2_; }
  const Point2& UnitDirection() const { return unit_direction_; }
  float Heading() const { return std::atan2(unit_direction_.y(), unit_direction_.x()); }

 private:
  Point2 p1_, p2_, unit_direction_;
  float length_;
};

// ---- include/ilqgames/geometry/polyline2.h:55-105 (point list + the route-walking accessors;
// closest-point queries run on the device)
class Polyline2 {
 public:
  Polyline2(const PointList2& points) : points_(points) {
    CHECK_GT(points.size(), 1u);
    for (size_t k = 1; k < points.size(); k++) segments_.emplace_back(points[k - 1], points[k]);
  }
  void AddPoint(const Point2& point) {
    segments_.emplace_back(points_.back(), point);
    points_.push_back(point);
  }
  const PointList2& Points() const { return points_; }
  const std::vector<LineSegment2>& Segments() const { return segments_; }
  size_t NumSegments() const { return segments_.size(); }
  float Length() const {
    float len = 0;
    for (const LineSegment2& seg : segments_) len += seg.Length();
    return len;
  }

  // The point `route_pos` metres along the polyline (src/polyline2.cpp:68-107): the segment that
  // contains that arclength (the last one when it lies beyond the end), walked from its first point.
  Point2 PointAt(float route_pos, bool* is_vertex = nullptr, LineSegment2* segment = nullptr,
                 bool* is_endpoint = nullptr) const {
    size_t idx = 0;
    float start = 0.0f;  // arclength at which segment idx begins
    while (idx + 1 < segments_.size() && start + segments_[idx].Length() <= route_pos) {
      start += segments_[idx].Length();
      idx++;
    }
    const LineSegment2& seg = segments_[idx];
    if (segment) *segment = seg;
    const float remaining = route_pos - start;
    CHECK_GE(remaining, 0.0f);
    if (is_vertex) *is_vertex = remaining < constants::kSmallNumber || remaining > seg.Length();
    const Point2 point = seg.FirstPoint() + remaining * seg.UnitDirection();
    if (is_endpoint)
      *is_endpoint = (idx == 0 || idx + 1 == segments_.size()) &&
                     (point == segments_.front().FirstPoint() || point == segments_.back().SecondPoint());
    return point;
  }

 private:
  PointList2 points_;
  std::vector<LineSegment2> segments_;
};

// Collects records and polylines while a Problem describes itself.
struct DescribeContext {
  ilqg_problem_desc* desc;
  int num_groups = 0;  // ExtremeValueCost objects described so far (their members share a group id)
  // identical point lists share one table entry (the four roundabout players reuse lanes)
  int AddPolyline(const Polyline2& polyline) {
    const PointList2& pts = polyline.Points();
    for (int p = 0; p < desc->num_polylines; p++) {
      const int s = desc->polyline_start[p], e = desc->polyline_start[p + 1];
      if (e - s != (int)pts.size()) continue;
      bool same = true;
      for (int k = 0; k < e - s && same; k++)
        same = desc->polyline_points[s + k][0] == pts[k].x() && desc->polyline_points[s + k][1] == pts[k].y();
      if (same) return p;
    }
    const int p = desc->num_polylines, s = desc->polyline_start[p];
    if (p >= ILQG_MAX_POLYLINES || s + (int)pts.size() > ILQG_MAX_POLYLINE_POINTS) return -1;
    for (size_t k = 0; k < pts.size(); k++) {
      desc->polyline_points[s + k][0] = pts[k].x();
      desc->polyline_points[s + k][1] = pts[k].y();
    }
    desc->polyline_start[p + 1] = s + (int)pts.size();
    desc->num_polylines = p + 1;
    return p;
  }
};

// ---- include/ilqgames/cost/cost.h:55-90, time_invariant_cost.h -------------------------------
class Cost {
 public:
  virtual ~Cost() {}
  void SetWeight(float weight) { weight_ = weight; }
  void ScaleWeight(float scale) { weight_ *= scale; }
  const std::string& Name() const { return name_; }
  float Weight() const { return weight_; }
  // fills kind / dim / flag / polyline / weight / value; the caller sets player, arg
  virtual bool Describe(ilqg_cost_desc* /*out*/, DescribeContext* /*ctx*/) const { return false; }
  // every record this cost contributes, in order: one for a plain cost, its members for a composite
  virtual bool DescribeAll(std::vector<ilqg_cost_desc>* out, DescribeContext* ctx) const {
    ilqg_cost_desc rec;
    std::memset(&rec, 0, sizeof(rec));
    rec.polyline = -1;
    if (!Describe(&rec, ctx)) return false;
    out->push_back(rec);
    return true;
  }

 protected:
  explicit Cost(float weight, const std::string& name = "") : name_(name), weight_(weight) {}
  const std::string name_;
  float weight_;
};

class TimeInvariantCost : public Cost {
 protected:
  explicit TimeInvariantCost(float weight, const std::string& name = "") : Cost(weight, name) {}
};

// src/quadratic_cost.cpp:51-94; dim < 0 = every dimension of the input
class QuadraticCost : public TimeInvariantCost {
 public:
  QuadraticCost(float weight, Dimension dim, float nominal = 0.0, const std::string& name = "")
      : TimeInvariantCost(weight, name), dimension_(dim), nominal_(nominal) {}
  bool Describe(ilqg_cost_desc* out, DescribeContext*) const override {
    out->kind = ILQG_COST_QUADRATIC;
    out->dim[0] = dimension_ < 0 ? -1 : dimension_;
    out->weight = weight_;
    out->value = nominal_;
    return true;
  }

 private:
  const Dimension dimension_;
  const float nominal_;
};

// src/quadratic_polyline2_cost.cpp:52-126
class QuadraticPolyline2Cost : public TimeInvariantCost {
 public:
  QuadraticPolyline2Cost(float weight, const Polyline2& polyline,
                         const std::pair<Dimension, Dimension>& position_idxs, const std::string& name = "")
      : TimeInvariantCost(weight, name), polyline_(polyline), xidx_(position_idxs.first), yidx_(position_idxs.second) {}
  bool Describe(ilqg_cost_desc* out, DescribeContext* ctx) const override {
    out->kind = ILQG_COST_QUADRATIC_POLYLINE2;
    out->dim[0] = xidx_;
    out->dim[1] = yidx_;
    out->weight = weight_;
    out->polyline = ctx->AddPolyline(polyline_);
    return out->polyline >= 0;
  }

 private:
  const Polyline2 polyline_;
  const Dimension xidx_, yidx_;
};

// src/proximity_cost.cpp:52-122
class ProximityCost : public TimeInvariantCost {
 public:
  ProximityCost(float weight, const std::pair<Dimension, Dimension>& position_idxs1,
                const std::pair<Dimension, Dimension>& position_idxs2, float threshold, const std::string& name = "")
      : TimeInvariantCost(weight, name), threshold_(threshold), xidx1_(position_idxs1.first),
        yidx1_(position_idxs1.second), xidx2_(position_idxs2.first), yidx2_(position_idxs2.second) {}
  bool Describe(ilqg_cost_desc* out, DescribeContext*) const override {
    out->kind = ILQG_COST_PROXIMITY;
    out->dim[0] = xidx1_; out->dim[1] = yidx1_; out->dim[2] = xidx2_; out->dim[3] = yidx2_;
    out->weight = weight_;
    out->value = threshold_;
    return true;
  }

 private:
  const float threshold_;
  const Dimension xidx1_, yidx1_, xidx2_, yidx2_;
};

// src/quadratic_difference_cost.cpp:50-91: half the weighted squared difference of two sets of
// dimensions (one or two pairs fit a record).  ILQG_COST_QUADRATIC_DIFFERENCE: CPU oracle only so far.
class QuadraticDifferenceCost : public TimeInvariantCost {
 public:
  QuadraticDifferenceCost(float weight, const std::vector<Dimension>& dims1, const std::vector<Dimension>& dims2,
                          const std::string& name = "")
      : TimeInvariantCost(weight, name), dims1_(dims1), dims2_(dims2) {
    CHECK_EQ(dims1_.size(), dims2_.size());
  }
  bool Describe(ilqg_cost_desc* out, DescribeContext*) const override {
    if (dims1_.empty() || dims1_.size() > 2) return false;
    out->kind = ILQG_COST_QUADRATIC_DIFFERENCE;
    for (size_t k = 0; k < dims1_.size(); k++) {
      out->dim[k] = dims1_[k];
      out->dim[2 + k] = dims2_[k];
    }
    out->flag = (int32_t)dims1_.size();
    out->weight = weight_;
    return true;
  }

 private:
  const std::vector<Dimension> dims1_, dims2_;
};

// src/signed_distance_cost.cpp:50-112 (include/ilqgames/cost/signed_distance_cost.h:52-88):
// nominal minus the distance between two positions, unweighted.  ILQG_COST_SIGNED_DISTANCE, which
// only the CPU oracle implements so far.  The defaulted bool sits before the name as in the
// reference, so a call that passes a string literal fourth binds it to less_is_positive there too.
class SignedDistanceCost : public TimeInvariantCost {
 public:
  SignedDistanceCost(const std::pair<Dimension, Dimension>& dims1, const std::pair<Dimension, Dimension>& dims2,
                     float nominal = 0.0, bool less_is_positive = true, const std::string& name = "")
      : TimeInvariantCost(1.0, name), xdim1_(dims1.first), ydim1_(dims1.second), xdim2_(dims2.first),
        ydim2_(dims2.second), nominal_(nominal), less_is_positive_(less_is_positive) {
    CHECK_GE(xdim1_, 0);
    CHECK_GE(ydim1_, 0);
    CHECK_GE(xdim2_, 0);
    CHECK_GE(ydim2_, 0);
  }
  bool Describe(ilqg_cost_desc* out, DescribeContext*) const override {
    out->kind = ILQG_COST_SIGNED_DISTANCE;
    out->dim[0] = xdim1_; out->dim[1] = ydim1_; out->dim[2] = xdim2_; out->dim[3] = ydim2_;
    out->weight = weight_;
    out->value = nominal_;
    out->flag = less_is_positive_;
    return true;
  }

 private:
  const Dimension xdim1_, ydim1_, xdim2_, ydim2_;
  const float nominal_;
  const bool less_is_positive_;
};

// src/semiquadratic_cost.cpp:51-85
class SemiquadraticCost : public TimeInvariantCost {
 public:
  SemiquadraticCost(float weight, Dimension dim, float threshold, bool oriented_right, const std::string& name = "")
      : TimeInvariantCost(weight, name), dimension_(dim), threshold_(threshold), oriented_right_(oriented_right) {
    CHECK_GE(dimension_, 0);
  }
  bool Describe(ilqg_cost_desc* out, DescribeContext*) const override {
    out->kind = ILQG_COST_SEMIQUADRATIC;
    out->dim[0] = dimension_;
    out->weight = weight_;
    out->value = threshold_;
    out->flag = oriented_right_;
    return true;
  }

 private:
  const Dimension dimension_;
  const float threshold_;
  const bool oriented_right_;
};

// src/semiquadratic_polyline2_cost.cpp:52-142
class SemiquadraticPolyline2Cost : public TimeInvariantCost {
 public:
  SemiquadraticPolyline2Cost(float weight, const Polyline2& polyline,
                             const std::pair<Dimension, Dimension>& position_idxs, float threshold,
                             bool oriented_right, const std::string& name = "")
      : TimeInvariantCost(weight, name), polyline_(polyline), xidx_(position_idxs.first),
        yidx_(position_idxs.second), threshold_(threshold), oriented_right_(oriented_right) {}
  bool Describe(ilqg_cost_desc* out, DescribeContext* ctx) const override {
    out->kind = ILQG_COST_SEMIQUADRATIC_POLYLINE2;
    out->dim[0] = xidx_;
    out->dim[1] = yidx_;
    out->weight = weight_;
    out->value = threshold_;
    out->flag = oriented_right_;
    out->polyline = ctx->AddPolyline(polyline_);
    return out->polyline >= 0;
  }

 private:
  const Polyline2 polyline_;
  const Dimension xidx_, yidx_;
  const float threshold_;
  const bool oriented_right_;
};

// src/polyline2_signed_distance_cost.cpp:52-121
class Polyline2SignedDistanceCost : public TimeInvariantCost {
 public:
  Polyline2SignedDistanceCost(const Polyline2& polyline, const std::pair<Dimension, Dimension>& position_idxs,
                              const float nominal = 0.0, bool oriented_same_as_polyline = true,
                              const std::string& name = "")
      : TimeInvariantCost(1.0, name), polyline_(polyline), xidx_(position_idxs.first), yidx_(position_idxs.second),
        nominal_(nominal), oriented_same_as_polyline_(oriented_same_as_polyline) {}
  bool Describe(ilqg_cost_desc* out, DescribeContext* ctx) const override {
    out->kind = ILQG_COST_POLYLINE2_SIGNED_DISTANCE;
    out->dim[0] = xidx_;
    out->dim[1] = yidx_;
    out->weight = weight_;
    out->value = nominal_;
    out->flag = oriented_same_as_polyline_;
    out->polyline = ctx->AddPolyline(polyline_);
    return out->polyline >= 0;
  }

 private:
  const Polyline2 polyline_;
  const Dimension xidx_, yidx_;
  const float nominal_;
  const bool oriented_same_as_polyline_;
};

// include/ilqgames/cost/final_time_cost.h:55-88: the wrapped cost counts only from threshold_time
// (relative to the initial time) on -- the wrapped cost's record with a time gate.
class FinalTimeCost : public Cost {
 public:
  ~FinalTimeCost() {}
  FinalTimeCost(const std::shared_ptr<const Cost>& cost, Time threshold_time, const std::string& name = "")
      : Cost(0.0, name), cost_(cost), threshold_time_(threshold_time) {
    CHECK_NOTNULL(cost.get());
  }
  bool Describe(ilqg_cost_desc* out, DescribeContext* ctx) const override {
    if (!cost_->Describe(out, ctx)) return false;
    out->active_from = threshold_time_;
    return true;
  }

 private:
  const std::shared_ptr<const Cost> cost_;
  const Time threshold_time_;
};

// include/ilqgames/cost/extreme_value_cost.h:54-80, src/extreme_value_cost.cpp:50-84: the largest
// (is_min: smallest) of several costs -- its members' records, tied together by a group id.  Only
// the CPU oracle implements groups so far.
class ExtremeValueCost : public Cost {
 public:
  ExtremeValueCost(const std::vector<std::shared_ptr<const Cost>>& costs, bool is_min, const std::string& name = "")
      : Cost(1.0, name), is_min_(is_min), costs_(costs) {}
  bool DescribeAll(std::vector<ilqg_cost_desc>* out, DescribeContext* ctx) const override {
    if (costs_.empty()) return false;
    const int group = ++ctx->num_groups;
    for (const auto& cost : costs_) {
      std::vector<ilqg_cost_desc> member;
      if (!cost->DescribeAll(&member, ctx) || member.size() != 1 || member[0].group != 0) return false;  // no nesting
      member[0].group = group;
      member[0].group_is_min = is_min_;
      out->push_back(member[0]);
    }
    return true;
  }

 private:
  const bool is_min_;
  const std::vector<std::shared_ptr<const Cost>> costs_;
};

// Costs the in-scope example sources include but never add to a player: constructible, not
// describable (ilqg_create would be refused).
#define ILQGAMES_B200_UNSUPPORTED_COST(Name)                                              \
  class Name : public TimeInvariantCost {                                                  \
   public:                                                                                 \
    template <typename... Args>                                                            \
    explicit Name(float weight, Args&&...) : TimeInvariantCost(weight, #Name) {}           \
  }
ILQGAMES_B200_UNSUPPORTED_COST(CurvatureCost);
ILQGAMES_B200_UNSUPPORTED_COST(LocallyConvexProximityCost);
ILQGAMES_B200_UNSUPPORTED_COST(NominalPathLengthCost);
ILQGAMES_B200_UNSUPPORTED_COST(OrientationCost);
ILQGAMES_B200_UNSUPPORTED_COST(WeightedConvexProximityCost);
#undef ILQGAMES_B200_UNSUPPORTED_COST

// ---- include/ilqgames/constraint/constraint.h:60-146 -----------------------------------------
// The multipliers (lambda per time step, mu) live in the device slab, per game; the reference
// keeps them in the object / in a process-wide static (src/constraint.cpp:61).
class Constraint {
 public:
  virtual ~Constraint() {}
  bool IsEquality() const { return is_equality_; }
  const std::string& Name() const { return name_; }
  virtual bool Describe(ilqg_cost_desc* /*out*/, DescribeContext* /*ctx*/) const { return false; }

 protected:
  explicit Constraint(bool is_equality, const std::string& name) : name_(name), is_equality_(is_equality) {}
  const std::string name_;
  const bool is_equality_;
};

class TimeInvariantConstraint : public Constraint {
 protected:
  explicit TimeInvariantConstraint(bool is_equality, const std::string& name) : Constraint(is_equality, name) {}
};

// src/proximity_constraint.cpp:56-116
class ProximityConstraint : public TimeInvariantConstraint {
 public:
  ProximityConstraint(const std::pair<Dimension, Dimension>& dims1, const std::pair<Dimension, Dimension>& dims2,
                      float threshold, bool keep_within, const std::string& name = "")
      : TimeInvariantConstraint(false, name), xidx1_(dims1.first), yidx1_(dims1.second), xidx2_(dims2.first),
        yidx2_(dims2.second), threshold_(threshold), keep_within_(keep_within) {
    CHECK_GT(threshold_, 0.0);
  }
  bool Describe(ilqg_cost_desc* out, DescribeContext*) const override {
    out->kind = ILQG_CONSTRAINT_PROXIMITY;
    out->dim[0] = xidx1_; out->dim[1] = yidx1_; out->dim[2] = xidx2_; out->dim[3] = yidx2_;
    out->value = threshold_;
    out->flag = keep_within_;
    out->weight = 1.0f;
    return true;
  }

 private:
  const Dimension xidx1_, yidx1_, xidx2_, yidx2_;
  const float threshold_;
  const bool keep_within_;
};

// include/ilqgames/constraint/single_dimension_constraint.h:62-96
class SingleDimensionConstraint : public TimeInvariantConstraint {
 public:
  SingleDimensionConstraint(Dimension dim, float threshold, bool keep_below, const std::string& name = "")
      : TimeInvariantConstraint(false, name), dim_(dim), threshold_(threshold), keep_below_(keep_below) {}
  bool Describe(ilqg_cost_desc* out, DescribeContext*) const override {
    out->kind = ILQG_CONSTRAINT_SINGLE_DIMENSION;
    out->dim[0] = dim_;
    out->value = threshold_;
    out->flag = keep_below_;
    out->weight = 1.0f;
    return true;
  }

 private:
  const Dimension dim_;
  const float threshold_;
  const bool keep_below_;
};

// include/ilqgames/constraint/polyline2_signed_distance_constraint.h:58-90,
// src/polyline2_signed_distance_constraint.cpp:58-145 (the intersection example constructs six of these
// -- its lane boundaries -- and adds none: src/three_player_intersection_example.cpp:214-251)
class Polyline2SignedDistanceConstraint : public TimeInvariantConstraint {
 public:
  Polyline2SignedDistanceConstraint(const Polyline2& polyline, const std::pair<Dimension, Dimension>& dims,
                                    float threshold, bool keep_left, const std::string& name = "")
      : TimeInvariantConstraint(false, name), polyline_(polyline), xidx_(dims.first), yidx_(dims.second),
        threshold_(threshold), keep_left_(keep_left) {}
  bool Describe(ilqg_cost_desc* out, DescribeContext* ctx) const override {
    out->kind = ILQG_CONSTRAINT_POLYLINE2_SIGNED_DISTANCE;
    out->dim[0] = xidx_;
    out->dim[1] = yidx_;
    out->value = threshold_;
    out->flag = keep_left_;
    out->weight = 1.0f;
    out->polyline = ctx->AddPolyline(polyline_);
    return out->polyline >= 0;
  }

 private:
  const Polyline2 polyline_;
  const Dimension xidx_, yidx_;
  const float threshold_;
  const bool keep_left_;
};

// ---- include/ilqgames/cost/player_cost.h:57-152 -----------------------------------------------
class PlayerCost {
 public:
  explicit PlayerCost(const std::string& name = "", float state_regularization = 0.0,
                      float control_regularization = 0.0)
      : name_(name), state_regularization_(state_regularization), control_regularization_(control_regularization),
        cost_structure_(ILQG_COST_SUM) {}

  void AddStateCost(const std::shared_ptr<Cost>& cost) { state_costs_.push_back(cost); }
  void AddControlCost(PlayerIndex idx, const std::shared_ptr<Cost>& cost) { control_costs_.emplace_back(idx, cost); }
  void AddStateConstraint(const std::shared_ptr<Constraint>& constraint) { state_constraints_.push_back(constraint); }
  void AddControlConstraint(PlayerIndex idx, const std::shared_ptr<Constraint>& constraint) {
    control_constraints_.emplace_back(idx, constraint);
  }

  void SetMaxOverTime() { cost_structure_ = ILQG_COST_MAX; }
  void SetMinOverTime() { cost_structure_ = ILQG_COST_MIN; }
  bool IsTimeAdditive() const { return cost_structure_ == ILQG_COST_SUM; }
  bool IsMaxOverTime() const { return cost_structure_ == ILQG_COST_MAX; }
  bool IsMinOverTime() const { return cost_structure_ == ILQG_COST_MIN; }
  int CostStructure() const { return cost_structure_; }

  const PtrVector<Cost>& StateCosts() const { return state_costs_; }
  const std::vector<std::pair<PlayerIndex, std::shared_ptr<Cost>>>& ControlCosts() const { return control_costs_; }
  const PtrVector<Constraint>& StateConstraints() const { return state_constraints_; }
  const std::vector<std::pair<PlayerIndex, std::shared_ptr<Constraint>>>& ControlConstraints() const {
    return control_constraints_;
  }
  bool IsConstrained() const { return !state_constraints_.empty() || !control_constraints_.empty(); }
  float StateRegularization() const { return state_regularization_; }
  float ControlRegularization() const { return control_regularization_; }
  const std::string& Name() const { return name_; }

 private:
  std::string name_;
  PtrVector<Cost> state_costs_;
  // the reference keeps control costs in an unordered_multimap keyed by player; insertion
  // order is what the records need (order-free for the in-scope examples)
  std::vector<std::pair<PlayerIndex, std::shared_ptr<Cost>>> control_costs_;
  PtrVector<Constraint> state_constraints_;
  std::vector<std::pair<PlayerIndex, std::shared_ptr<Constraint>>> control_constraints_;
  float state_regularization_, control_regularization_;
  int cost_structure_;
};

}  // namespace ilqgames

#endif
