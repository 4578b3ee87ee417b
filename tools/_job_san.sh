for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool tools/sanitize_small.py"
  compute-sanitizer --tool $tool python tools/sanitize_small.py 2>&1 | grep -E "SUMMARY|sanitize run ok|Error|error" | head -8
  echo "== compute-sanitizer --tool $tool tools/sanitize_ring.py"
  compute-sanitizer --tool $tool python tools/sanitize_ring.py 2>&1 | grep -E "SUMMARY|sanitize run ok|Error|error" | head -8
done
