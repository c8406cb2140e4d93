# round 2, check 3: full GPU test tier after the BART / API / kernel additions
python -m pytest tests -m gpu -q 2>&1 | tail -40
