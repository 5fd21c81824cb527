# Builds libgvit.so (the C-ABI CUDA library) for sm_100a.  `make -j` compiles the translation units in parallel.
NVCC      ?= nvcc
PKG       := graph_augmented_vision_transformers_b200
SRC_DIR   := $(PKG)/csrc
OBJ_DIR   := build/obj
LIB       := $(PKG)/lib/libgvit.so
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden
SRCS      := $(wildcard $(SRC_DIR)/*.cu)
OBJS      := $(patsubst $(SRC_DIR)/%.cu,$(OBJ_DIR)/%.o,$(SRCS))
HDRS      := $(wildcard $(SRC_DIR)/*.cuh) include/gvit.h

all: $(LIB)

$(OBJ_DIR)/%.o: $(SRC_DIR)/%.cu $(HDRS)
	@mkdir -p $(OBJ_DIR)
	$(NVCC) $(NVCCFLAGS) $(EXTRA) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p $(dir $(LIB))
	$(NVCC) -shared -o $@ $(OBJS) -cudart static

clean:
	rm -rf build $(LIB)

.PHONY: all clean
