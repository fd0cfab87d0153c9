python tools/copyprobe.py 4097
B2_NOBULK1D=1 python tools/copyprobe.py 4097
B2_NOTMA=1 python tools/copyprobe.py 4097
B2_LN=2 python tools/copyprobe.py 4097
python tools/copyprobe.py 1025
