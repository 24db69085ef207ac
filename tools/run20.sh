python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/run_mine.py --config 4 --repeat 2 | cut -c1-200
python tools/run_mine.py --config 1 --repeat 3 | cut -c1-300
python tools/run_mine.py --config 2 --repeat 2 | cut -c1-200
