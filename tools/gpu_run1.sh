bash tools/prof_one.sh BC7 bc7_encode bc7_r2g
bash tools/prof_one.sh ETC2_RGBA etc_encode etc2_r2g
