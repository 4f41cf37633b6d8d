// The reference includes "Slam.h" (EKF.h:3, PF.h:3) but ships slam.h — a case-insensitive
// file-system assumption.  Forward to the real header where it lies.
#pragma once
#include "slam.h"
