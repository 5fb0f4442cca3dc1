// Forwarding header: lets the reference's autoencoder.cpp keep its `#include "netlib.h"` line unchanged.
#include "aefft_shim.h"
