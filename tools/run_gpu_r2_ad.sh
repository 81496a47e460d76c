# Round 2, GPU call AD: last sanity of the committed build (training + model ABI + parity suites, smoke)
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3 | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
