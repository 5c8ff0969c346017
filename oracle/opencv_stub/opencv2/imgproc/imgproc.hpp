// see opencv2/core/core.hpp in this directory (stand-in headers, test infrastructure only)
#include "opencv2/core/core.hpp"
