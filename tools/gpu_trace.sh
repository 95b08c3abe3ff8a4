#!/bin/bash
for d in 0 7 1; do for role in 1; do echo "== dbg $d role $role"; SMB_WS_DBG=$((16 + d + role * 256)) python tools/ws_trace.py 2>&1 | tail -14; done; done
