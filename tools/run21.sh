python tools/run_mine.py --config 4 --repeat 2 | cut -c1-200
python tools/run_mine.py --config 3 --repeat 1 | cut -c1-200
python tools/run_mine.py --config 5 --repeat 1 | cut -c1-200
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
