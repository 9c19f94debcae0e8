"""The MoveIt-side boundary meets a compiler.  moveit2_b200/adapter/ (the CollisionEnv subclass, the allocator, the pluginlib
loader, the OMPL validators) needs MoveIt 2, Eigen, geometric_shapes, rclcpp, pluginlib and OMPL headers, none of which
this image has; tests/mock_moveit/ declares the members it uses with the reference's signatures (each block cites the
reference header), and every adapter source must compile against it with -Wall -Wextra -Werror.  Declarations only:
nothing is linked or run, so this checks names, argument types, const-correctness and that every pure virtual of
collision_detection::CollisionEnv is overridden (the allocator template instantiates CollisionEnvB200)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "moveit2_b200", "adapter")
FLAGS = ["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-c", "-I", os.path.join(ROOT, "tests", "mock_moveit"),
         "-I", os.path.join(ROOT, "include"), "-I", ADAPTER]


@pytest.mark.parametrize("source", ["collision_env_b200.cpp", "collision_detector_b200_plugin_loader.cpp"])
def test_adapter_source_compiles(tmp_path, source):
    out = subprocess.run(FLAGS + ["-o", str(tmp_path / (source + ".o")), os.path.join(ADAPTER, source)],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_ompl_validators_header_compiles_and_is_instantiable(tmp_path):
    src = tmp_path / "use_validators.cpp"
    src.write_text('#include "ompl_validators_b200.hpp"\n'
                   "// both validators must be concrete (every pure virtual of the OMPL bases overridden)\n"
                   "ompl::base::MotionValidator* mv(const ompl::base::SpaceInformationPtr& si,\n"
                   "                                const ompl_interface::ModelBasedPlanningContext* pc, b200sv_ctx* c)\n"
                   "{ return new b200::B200MotionValidator(si, pc, c); }\n"
                   "ompl::base::StateValidityChecker* sv(const ompl::base::SpaceInformationPtr& si, b200sv_ctx* c)\n"
                   "{ return new b200::B200StateValidityChecker(si, c); }\n")
    out = subprocess.run(FLAGS + ["-o", str(tmp_path / "use_validators.o"), str(src)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_allocator_instantiates_every_constructor(tmp_path):
    """CollisionDetectorAllocatorTemplate calls the three constructors CollisionEnv subclasses must have
    (collision_detector_allocator.hpp:80-93) and needs CollisionEnvB200 to be concrete."""
    src = tmp_path / "use_allocator.cpp"
    src.write_text('#include "collision_detector_allocator_b200.hpp"\n'
                   "collision_detection::CollisionEnvPtr a(const collision_detection::CollisionDetectorAllocatorPtr& al,\n"
                   "    const collision_detection::WorldPtr& w, const moveit::core::RobotModelConstPtr& m)\n"
                   "{ auto e = al->allocateEnv(w, m); auto f = al->allocateEnv(e, w); auto g = al->allocateEnv(m);\n"
                   "  return f ? f : g; }\n"
                   "collision_detection::CollisionDetectorAllocatorPtr b()\n"
                   "{ return collision_detection::CollisionDetectorAllocatorB200::create(); }\n")
    out = subprocess.run(FLAGS + ["-o", str(tmp_path / "use_allocator.o"), str(src)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_a_broken_adapter_would_not_compile(tmp_path):
    """The mock is strict enough to catch the mistakes it is there for: calling a member that does not exist, and dropping
    a pure-virtual override."""
    bad = tmp_path / "bad.cpp"
    bad.write_text('#include "collision_env_b200.hpp"\n'
                   "double f(const moveit::core::RobotState& s) { return s.getVariablePositionz(0); }\n")
    assert subprocess.run(FLAGS + ["-o", str(tmp_path / "bad.o"), str(bad)], capture_output=True).returncode != 0
    bad.write_text('#include <moveit/collision_detection/collision_env.hpp>\n'
                   "struct E : collision_detection::CollisionEnv { using CollisionEnv::CollisionEnv; };\n"
                   "collision_detection::CollisionEnv* g(const moveit::core::RobotModelConstPtr& m) { return new E(m); }\n")
    assert subprocess.run(FLAGS + ["-o", str(tmp_path / "bad2.o"), str(bad)], capture_output=True).returncode != 0
