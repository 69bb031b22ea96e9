#include "device_scan.cuh"
