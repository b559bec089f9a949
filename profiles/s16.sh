set -u
for inner in 2 3 4 6; do
echo "== svd D=4096 decay inner $inner"
QTB_SVD_INNER=$inner QTB_SVD_DEBUG=1 timeout 300 python profiles/svd_driver.py 15 4096 1.6 decay 2>&1 | tail -2
done
for inner in 2 3; do
echo "== svd D=4096 random inner $inner"
QTB_SVD_INNER=$inner QTB_SVD_DEBUG=1 timeout 300 python profiles/svd_driver.py 15 4096 1.6 2>&1 | tail -2
echo "== svd D=1024 decay inner $inner"
QTB_SVD_INNER=$inner QTB_SVD_DEBUG=1 timeout 300 python profiles/svd_driver.py 15 1024 1.6 decay 2>&1 | tail -2
done
