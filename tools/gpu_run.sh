timeout 200 python tools/read_bw.py
