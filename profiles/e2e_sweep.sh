for slots in 3 4 6; do for chunk in 524288 1048576 2097152; do GB_EVAL_SLOTS=$slots GB_EVAL_CHUNK=$chunk python profiles/time_e2e.py 2>&1 | grep "tiled=None" | sed "s/^/slots=$slots /"; done; done
python profiles/time_e2e.py 2>&1 | tail -3
