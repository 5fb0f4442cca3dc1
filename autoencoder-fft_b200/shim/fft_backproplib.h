// Forwarding header: lets the reference's autoencoder.cpp keep its `#include "fft_backproplib.h"` line unchanged.
#include "aefft_shim.h"
