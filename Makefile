# Top-level convenience targets; the driver uses __graft_entry__.build().
all:
	$(MAKE) -C paf_baseband2power_b200/csrc all
	$(MAKE) -C paf_baseband2power_b200/host all
	$(MAKE) -C oracle all

test-cpu: all
	python -m pytest tests -x -q -m "not gpu"

test-gpu: all
	python -m pytest tests -x -q -m gpu

bench: all
	python bench.py

clean:
	$(MAKE) -C paf_baseband2power_b200/csrc clean
	$(MAKE) -C paf_baseband2power_b200/host clean
	$(MAKE) -C oracle clean

.PHONY: all test-cpu test-gpu bench clean
