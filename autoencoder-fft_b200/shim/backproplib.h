// Forwarding header: lets the reference's autoencoder.cpp keep its `#include "backproplib.h"` line unchanged.
#include "aefft_shim.h"
