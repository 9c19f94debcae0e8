// OMPL-side consumers of the batched C-ABI: a MotionValidator and a StateValidityChecker for MoveIt's OMPL interface.
// Header only; compiled where MoveIt 2 and OMPL are installed (not in this repository's image).  INTEGRATION.md 3b.
//
// What they replace in the reference:
//   * MoveIt installs no MotionValidator: OMPL's DiscreteMotionValidator interpolates an edge into
//     StateSpace::validSegmentCount states and calls StateValidityChecker::isValid once per state
//     (moveit_planners/ompl/ompl_interface/src/detail/state_validity_checker.cpp:83-133), each call copying the state into
//     a RobotState, updating its transforms and running PlanningScene::isStateColliding.
//   * Here an edge (or M edges) is ONE b200sv_check_motions call: interpolation, FK and the checks run on the device.
// Install next to setStateValidityChecker in ModelBasedPlanningContext::configure
// (model_based_planning_context.cpp:110-140):
//     si->setMotionValidator(std::make_shared<b200::B200MotionValidator>(si, this, ctx));
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <utility>
#include <vector>

#include <moveit/ompl_interface/model_based_planning_context.hpp>
#include <moveit/ompl_interface/parameterization/model_based_state_space.hpp>
#include <moveit/robot_model/revolute_joint_model.hpp>
#include <ompl/base/MotionValidator.h>
#include <ompl/base/SpaceInformation.h>
#include <ompl/base/StateValidityChecker.h>

#include "b200sv.h"

namespace b200
{
using ompl_interface::ModelBasedPlanningContext;
using ompl_interface::ModelBasedStateSpace;

// Per group variable: continuous flag and JointModel::getDistanceFactor (0 for the variables of mimic joints, which
// JointModelGroup::distance skips: joint_model_group.cpp:462-472 iterates the ACTIVE joints).
struct GroupMetric
{
  std::vector<uint8_t> continuous;
  std::vector<double> factor;
  bool one_dof_joints_only = true;  // planar / floating joints: b200sv_check_motions refuses them (their metric is not
                                    // per-variable); the validators then fall back to state-by-state b200sv_check_batch
  explicit GroupMetric(const moveit::core::JointModelGroup* jmg)
  {
    for (const moveit::core::JointModel* jm : jmg->getJointModels())
      for (std::size_t v = 0; v < jm->getVariableCount(); ++v)
      {
        if (jm->getType() == moveit::core::JointModel::PLANAR || jm->getType() == moveit::core::JointModel::FLOATING)
          one_dof_joints_only = false;
        continuous.push_back(jm->getType() == moveit::core::JointModel::REVOLUTE &&
                             static_cast<const moveit::core::RevoluteJointModel*>(jm)->isContinuous());
        factor.push_back(jm->getMimic() ? 0.0 : jm->getDistanceFactor());
      }
  }
};

class B200MotionValidator : public ompl::base::MotionValidator
{
public:
  B200MotionValidator(const ompl::base::SpaceInformationPtr& si, const ModelBasedPlanningContext* pc, b200sv_ctx* ctx)
    : ompl::base::MotionValidator(si), ctx_(ctx), metric_(pc->getJointModelGroup())
  {
    // longest_valid_segment_fraction * getMaximumExtent (model_based_planning_context.cpp:294-317)
    max_segment_ = si->getStateSpace()->getLongestValidSegmentLength();
    n_vars_ = pc->getJointModelGroup()->getVariableCount();
    const ompl::base::StateSpacePtr space = si->getStateSpace();
    distance_ = [space](const ompl::base::State* a, const ompl::base::State* b) { return space->distance(a, b); };
  }

  // `s1` is assumed valid, as OMPL's DiscreteMotionValidator assumes
  bool checkMotion(const ompl::base::State* s1, const ompl::base::State* s2) const override
  {
    if (!metric_.one_dof_joints_only)
    {
      std::pair<ompl::base::State*, double> ignored(nullptr, 0.0);
      return checkMotion(s1, s2, ignored);
    }
    uint8_t bit = 0;
    const bool ok = b200sv_check_motions(ctx_, values(s1), values(s2), 1, max_segment_, metric_.continuous.data(),
                                         metric_.factor.data(), 0, B200SV_CHECK_ALL, &bit, nullptr, nullptr) == B200SV_OK &&
                    (bit & 1);
    (ok ? valid_ : invalid_)++;
    return ok;
  }

  // lastValid: the last state before the first colliding one and its parameter along the edge
  // (DiscreteMotionValidator: lastValid.second = (j - 1) / nd, state = interpolate(s1, s2, (j - 1) / nd))
  bool checkMotion(const ompl::base::State* s1, const ompl::base::State* s2,
                   std::pair<ompl::base::State*, double>& last_valid) const override
  {
    if (!metric_.one_dof_joints_only)
      return checkMotionStateByState(s1, s2, last_valid);
    uint8_t bit = 0;
    int32_t first_invalid = -1;
    uint64_t nd = 0;
    const int rc = b200sv_check_motions(ctx_, values(s1), values(s2), 1, max_segment_, metric_.continuous.data(),
                                        metric_.factor.data(), 0, B200SV_CHECK_ALL, &bit, &first_invalid, &nd);
    const bool ok = rc == B200SV_OK && (bit & 1);
    if (!ok)
    {
      last_valid.second = (rc == B200SV_OK && nd > 0 && first_invalid > 0) ? static_cast<double>(first_invalid - 1) / nd : 0.0;
      if (last_valid.first)
        si_->getStateSpace()->interpolate(s1, s2, last_valid.second, last_valid.first);
    }
    (ok ? valid_ : invalid_)++;
    return ok;
  }

  // The batched entry: planners that grow several edges per iteration (batch RRTConnect, PRM roadmap construction, lazy
  // planners validating a whole path) validate M edges with one launch sequence.  from / to: [M x n_group_vars] doubles
  // in ModelBasedStateSpace::StateType::values layout; valid_bits: ceil(M / 8) bytes; first_invalid: [M] or nullptr.
  bool checkMotions(const double* from, const double* to, std::size_t m, uint8_t* valid_bits, int32_t* first_invalid) const
  {
    return b200sv_check_motions(ctx_, from, to, m, max_segment_, metric_.continuous.data(), metric_.factor.data(), 0,
                                B200SV_CHECK_ALL, valid_bits, first_invalid, nullptr) == B200SV_OK;
  }

private:
  static const double* values(const ompl::base::State* s)
  {
    return s->as<ModelBasedStateSpace::StateType>()->values;  // the layout b200sv expects: no copy
  }
  // Groups with planar / floating joints: the state space interpolates (slerp etc.), the device checks the nd states of
  // the edge as ONE batch.  Fails closed: any error reads as "edge invalid".
  bool checkMotionStateByState(const ompl::base::State* s1, const ompl::base::State* s2,
                               std::pair<ompl::base::State*, double>& last_valid) const
  {
    const double len = distance_ ? distance_(s1, s2) : 0.0;
    const std::size_t nd = std::max<std::size_t>(1, static_cast<std::size_t>(std::ceil(len / max_segment_)));
    std::vector<double> q(nd * n_vars_);
    std::vector<ModelBasedStateSpace::StateType> scratch(1);
    std::vector<double> tmp(n_vars_);
    scratch[0].values = tmp.data();
    for (std::size_t j = 1; j <= nd; ++j)
    {
      si_->getStateSpace()->interpolate(s1, s2, static_cast<double>(j) / nd, &scratch[0]);
      std::copy(tmp.begin(), tmp.end(), q.begin() + (j - 1) * n_vars_);
    }
    std::vector<uint8_t> bits((nd + 7) / 8, 0);
    const bool called = b200sv_check_batch(ctx_, q.data(), nd, B200SV_CHECK_ALL, bits.data()) == B200SV_OK;
    std::size_t first_bad = nd;
    for (std::size_t j = 0; j < nd && first_bad == nd; ++j)
      if (!called || !((bits[j >> 3] >> (j & 7)) & 1))
        first_bad = j;
    const bool ok = first_bad == nd;
    if (!ok)
    {
      last_valid.second = static_cast<double>(first_bad) / nd;
      if (last_valid.first)
        si_->getStateSpace()->interpolate(s1, s2, last_valid.second, last_valid.first);
    }
    (ok ? valid_ : invalid_)++;
    return ok;
  }
  b200sv_ctx* ctx_;
  GroupMetric metric_;
  double max_segment_;
  std::size_t n_vars_ = 0;
  std::function<double(const ompl::base::State*, const ompl::base::State*)> distance_;  // StateSpace::distance
};

// StateValidityChecker::isValid over b200sv_check_batch (a batch of one on the latency path), plus the batched form.
// Bounds and path constraints stay with the caller, as in ompl_interface::StateValidityChecker::isValid
// (state_validity_checker.cpp:83-133): this class answers the collision part only.
class B200StateValidityChecker : public ompl::base::StateValidityChecker
{
public:
  B200StateValidityChecker(const ompl::base::SpaceInformationPtr& si, b200sv_ctx* ctx)
    : ompl::base::StateValidityChecker(si), ctx_(ctx)
  {
  }

  bool isValid(const ompl::base::State* state) const override
  {
    if (!si_->satisfiesBounds(state))
      return false;
    uint8_t bit = 0;
    return b200sv_check_batch(ctx_, state->as<ModelBasedStateSpace::StateType>()->values, 1, B200SV_CHECK_ALL, &bit) ==
               B200SV_OK &&
           (bit & 1);
  }

  // n states, [n x n_group_vars] doubles; valid_bits: ceil(n / 8) bytes
  bool isValidBatch(const double* values, std::size_t n, uint8_t* valid_bits) const
  {
    return b200sv_check_batch(ctx_, values, n, B200SV_CHECK_ALL, valid_bits) == B200SV_OK;
  }

private:
  b200sv_ctx* ctx_;
};
}  // namespace b200
