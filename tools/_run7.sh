set -x
timeout 900 python tools/check_attn_ps.py 2>&1 | tail -14
