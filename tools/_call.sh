python tools/time_trainer.py 128 40 2>/dev/null
