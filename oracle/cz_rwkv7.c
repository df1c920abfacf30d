/* placeholder until RWKV-7 lands */
#include "cz_oracle.h"
czo_session *czo_rwkv7_new(const czo_rwkv7_config *cfg) { (void)cfg; return 0; }
