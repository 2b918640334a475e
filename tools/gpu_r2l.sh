for r in 1 2; do
CNH_E2E_STAGE_FIRST=1 timeout 300 python tools/e2e_legs.py boxes boxes:eager | sed 's/^/stage-first /'
CNH_E2E_STAGE_FIRST=0 timeout 300 python tools/e2e_legs.py boxes boxes:eager | sed 's/^/run-first   /'
done
