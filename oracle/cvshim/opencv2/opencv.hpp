// oracle/cvshim — test infrastructure, see opencv2/core/core.hpp
#pragma once
#include "opencv2/core/core.hpp"
#include "opencv2/features2d/features2d.hpp"
#include "opencv2/imgproc/imgproc.hpp"
