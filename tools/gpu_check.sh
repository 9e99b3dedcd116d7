cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/prof_pretok.py tinystories 1000000000 2>&1 | tail -3
timeout 300 python tools/prof_pretok.py owt 1000000000 2>&1 | tail -3
