python tools/e2e_pageable.py
for t in 4 8 16; do ILCC2_STAGE_THREADS=$t python tools/e2e_pageable.py; done
ILCC2_DRIVER_STAGING=1 python tools/e2e_pageable.py
ILCC2_E2E_CHUNK=64 python tools/e2e_pageable.py
ILCC2_E2E_CHUNK=256 python tools/e2e_pageable.py
