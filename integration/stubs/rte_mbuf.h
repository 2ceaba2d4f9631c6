#include "rte_bbdev.h"
