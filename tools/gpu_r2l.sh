timeout 300 python tools/e2e_legs.py
