python scripts/profile_step.py --steps 6 | tail -3
