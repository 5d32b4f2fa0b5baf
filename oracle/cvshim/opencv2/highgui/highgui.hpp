// oracle/cvshim — test infrastructure, see opencv2/core/core.hpp (nothing of highgui is used on the path)
#pragma once
#include "opencv2/core/core.hpp"
