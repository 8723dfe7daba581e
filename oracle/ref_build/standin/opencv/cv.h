// TEST INFRASTRUCTURE: stand-in, see opencv2/core/core.hpp
#include "opencv2/core/core.hpp"
