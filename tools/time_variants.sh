#!/bin/bash
# Development aid (GPU box): time the default library and every tools/libosudit_v_*.so with tools/attn_time.py.
cd "$(dirname "$0")/.."
timeout 60 python tools/attn_time.py stream 2>&1 | grep -v Warning
for so in tools/libosudit_v_*.so; do [ -e "$so" ] && OSUDIT_LIB=$PWD/$so timeout 60 python tools/attn_time.py stream 2>&1 | grep -v Warning; done
