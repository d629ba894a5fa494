echo "== HEAD"; python tools/r04t.py
echo "== no unit codes"; MGQP_NO_UNIT_CODES=1 python tools/r04t.py
echo "== stream sync (old)"; MGQP_HOST_STREAM_SYNC=1 MGQP_NO_UNIT_CODES=1 python tools/r04t.py
