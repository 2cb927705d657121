timeout 200 python scripts/status_diag.py stag 8192 1100 "[128,384]"
timeout 200 python scripts/status_diag.py stag 8192 1100 "[224,416]"
timeout 200 python scripts/status_diag.py eco 16384 600 "[224,416]" --rich
timeout 200 python scripts/status_diag.py base 4096 1500
timeout 200 python scripts/status_diag.py eco 16384 1500
