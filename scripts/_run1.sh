python scripts/e2e_pipeline.py 12 1 2 3 4 6
python scripts/e2e_phases.py | tail -3
