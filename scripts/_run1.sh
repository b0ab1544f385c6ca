python scripts/profile_step.py --steps 6 | tail -2
python scripts/rows_variants.py
