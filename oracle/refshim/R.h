#pragma once
// stand-in for R.h (oracle/refshim): nothing needed beyond RcppArmadillo.h
#include "RcppArmadillo.h"
