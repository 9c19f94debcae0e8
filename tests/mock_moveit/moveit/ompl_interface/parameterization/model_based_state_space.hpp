// forwards to the one mock header (tests/mock_moveit/mock_moveit_all.hpp): test infrastructure
#pragma once
#include "mock_moveit_all.hpp"
