FQB200_LIB=$PWD/tools/_san/libfqb200_loads_only.so python tools/scan_skeleton.py 0 1 2 4 5 6
python tools/scan_skeleton.py 0
