( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
