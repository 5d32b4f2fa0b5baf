// oracle/refstub — see orb_slam3_stubs.h (Sophus::SE3f / Sim3f live there)
#pragma once
#include "../orb_slam3_stubs.h"
