/* Source-compatibility forwarder: the reference's <bifrost/ring.h>.
 * All declarations live in the consolidated <bifrost_b200.h>. */
#pragma once
#include "../bifrost_b200.h"
