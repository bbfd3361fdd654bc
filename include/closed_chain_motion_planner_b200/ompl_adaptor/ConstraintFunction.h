// Drop-in replacement of the reference's
//   include/closed_chain_motion_planner/base/constraints/ConstraintFunction.h   (ConstraintFunction.h:21-137)
// for a tree that has OMPL and Eigen: the same class name, base class, virtuals and helper methods, so that
// main.cpp:41, ConstrainedPlanningCommon.cpp:126-129 and jy_ProjectedStateSpace.cpp:13,20,27,65 compile unchanged — the
// arithmetic runs in libccp.so (include/ccp.h).  It uses only the parts of Eigen::Ref / Matrix that are stable across
// Eigen 3.3 / 3.4: data(), size(), operator[] and operator()(i, j).
//
// OMPL and Eigen are not installed in this repository's image: tests/test_ompl_adaptor_gpu.py compiles this header
// against minimal stand-ins of the two interfaces (tests/stubs/) and drives it through ompl::base::Constraint's virtuals
// on the GPU — a syntax and behaviour check of every signature cited here, not a test of OMPL.
#pragma once
#include <ompl/base/Constraint.h>
#include <ompl/base/spaces/constraint/ConstrainedStateSpace.h>
#include <ompl/util/Exception.h>

#include <closed_chain_motion_planner/kinematics/panda_model.h>

#include <algorithm>
#include <cstdint>
#include <memory>
#include <vector>

#include "ccp.h"

class KinematicChainConstraint : public ompl::base::Constraint
{
public:
    KinematicChainConstraint(unsigned int links) : ompl::base::Constraint(links, 2 * (links / 7 - 1))  // ConstraintFunction.h:24
    {
    }
    ~KinematicChainConstraint() override
    {
        ccp_host_free(buf_);
        ccp_destroy(h_);
    }

    void setArmModels(const ArmModelPtr &arm1, const ArmModelPtr &arm2)  // ConstraintFunction.h:122-126
    {
        ccp_model_desc d;
        int32_t idx[2] = {arm1->index, arm2->index};
        if (ccp_default_model(2, idx, &d) != CCP_OK)  // stock DH, limits, flange, yaw
            throw ompl::Exception("KinematicChainConstraint: bad arm index");
        const ArmModelPtr arms[2] = {arm1, arm2};
        for (int a = 0; a < 2; ++a)  // t_wb (grasping_point.cpp:11-16) as row-major 3x4
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 4; ++c)
                    d.arm[a].t_wb[4 * r + c] = arms[a]->t_wb.matrix()(r, c);
        // calibrated arms: add the 7x4 offsets (a, d, theta, alpha) into d.arm[a].dh_* (panda_rbdl.cpp:92-95)
        ccp_destroy(h_);
        h_ = nullptr;
        if (ccp_create(&d, /*device*/ 0, &h_) != CCP_OK)
            throw ompl::Exception(ccp_last_error(nullptr));
    }

    void setInitialPosition(const Eigen::Ref<const Eigen::VectorXd> init_joint)  // ConstraintFunction.h:31-40
    {
        check(ccp_set_reference(need(), init_joint.data()));
    }

    void setTolerance(const double tolerance1, const double tolerance2)  // ConstraintFunction.h:104-112
    {
        if (tolerance1 <= 0 || tolerance2 <= 0)
            throw ompl::Exception("ompl::base::Constraint::setProjectionTolerance(): tolerance must be positive.");
        tolerance1_ = tolerance1;
        tolerance2_ = tolerance2;
        check(ccp_set_tolerance(need(), tolerance1, tolerance2));
    }

    void function(const Eigen::Ref<const Eigen::VectorXd> &x, Eigen::Ref<Eigen::VectorXd> out) const override  // :84-102
    {
        double f[4];
        check(ccp_function_batch_host(need(), x.data(), 1, f));
        for (unsigned int i = 0; i < getCoDimension(); ++i)
            out[i] = f[i];
    }

    // the reference inherits OMPL's finite-difference default (called at :70); here it is analytic
    void jacobian(const Eigen::Ref<const Eigen::VectorXd> &x, Eigen::Ref<Eigen::MatrixXd> out) const override
    {
        std::vector<double> J((size_t)getCoDimension() * getAmbientDimension());  // row-major m x n
        check(ccp_jacobian_batch_host(need(), x.data(), 1, J.data()));
        for (unsigned int i = 0; i < getCoDimension(); ++i)
            for (unsigned int j = 0; j < getAmbientDimension(); ++j)
                out(i, j) = J[(size_t)i * getAmbientDimension() + j];
    }

    bool project(Eigen::Ref<Eigen::VectorXd> x) const override  // ConstraintFunction.h:57-82
    {
        uint8_t ok = 0;  // x is updated in place, the last iterate also on failure
        check(ccp_project_batch_host(need(), x.data(), 1, x.data(), &ok, nullptr, nullptr, nullptr));
        return ok != 0;
    }
    using ompl::base::Constraint::project;  // project(State *) used by the samplers (jy_ProjectedStateSpace.cpp:13,20,27,65)

    bool isSatisfied(const Eigen::Ref<const Eigen::VectorXd> &x) const override  // ConstraintFunction.h:114-120
    {
        double f[4];
        check(ccp_function_batch_host(need(), x.data(), 1, f));
        return f[0] - f[0] == 0.0 && f[1] - f[1] == 0.0 && f[0] <= tolerance1_ && f[1] <= tolerance2_;
    }

    bool jointValid(const Eigen::Ref<const Eigen::VectorXd> &q) const  // ConstraintFunction.h:43-55
    {
        static const double lb[7] = {-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973};
        static const double ub[7] = {2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973};
        for (int arm = 0; arm < 2; ++arm)
            for (int i = 0; i < 7; ++i)
                if (q[arm * 7 + i] < lb[i] + 1e-3 || q[arm * 7 + i] > ub[i] - 1e-3)
                    return false;
        return true;
    }

    // NEW (north star): batched projection of gathered OMPL states; ok[i] = project()'s return value for states[i]
    void projectBatch(const std::vector<ompl::base::State *> &states, std::vector<uint8_t> &ok) const
    {
        const size_t c = states.size(), n = getAmbientDimension();
        if (c * n > buf_len_)  // page-locked staging, allocated once: the chunked copies then overlap the kernels
        {
            ccp_host_free(buf_);
            buf_ = nullptr;
            if (ccp_host_alloc((void **)&buf_, sizeof(double) * c * n) != CCP_OK)
                throw ompl::Exception("projectBatch: ccp_host_alloc failed");
            buf_len_ = c * n;
        }
        for (size_t i = 0; i < c; ++i)  // gather (the states are individually allocated: KinematicChain.cpp:97)
            std::copy_n(states[i]->as<ompl::base::ConstrainedStateSpace::StateType>()->data(), n, buf_ + i * n);
        ok.resize(c);
        check(ccp_project_batch_host(need(), buf_, (int64_t)c, buf_, ok.data(), nullptr, nullptr, nullptr));
        for (size_t i = 0; i < c; ++i)  // scatter back, failures too
            std::copy_n(buf_ + i * n, n, states[i]->as<ompl::base::ConstrainedStateSpace::StateType>()->data());
    }

    ccp_handle *handle() const { return h_; }

private:
    ccp_handle *need() const
    {
        if (!h_)
            throw ompl::Exception("KinematicChainConstraint: setArmModels() must be called first");
        return h_;
    }
    void check(int rc) const
    {
        if (rc != CCP_OK)
            throw ompl::Exception(ccp_last_error(h_));
    }
    ccp_handle *h_ = nullptr;
    mutable double *buf_ = nullptr;
    mutable size_t buf_len_ = 0;
    double tolerance1_ = 1e-3, tolerance2_ = 5e-3;  // ConstrainedPlanningCommon.cpp:120-121
};
typedef std::shared_ptr<KinematicChainConstraint> ChainConstraintPtr;  // ConstraintFunction.h:140
