bash tools/prof_one.sh BC6HU bc6h_encode bc6hu_r2k
