# Builds libsea_b200.so (sm_100a only) and the oracle helpers.  `python -c "import __graft_entry__ as g; g.build()"`
# drives the same commands.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC
SRCS := $(wildcard sea_b200/csrc/*.cu)
OBJS := $(patsubst sea_b200/csrc/%.cu,build/%.o,$(SRCS))
LIB := sea_b200/lib/libsea_b200.so

all: $(LIB)

build/%.o: sea_b200/csrc/%.cu $(wildcard sea_b200/csrc/*.h sea_b200/csrc/*.cuh) include/sea_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p sea_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart_static -lpthread -ldl -lrt

clean:
	rm -rf build $(LIB)
.PHONY: all clean
