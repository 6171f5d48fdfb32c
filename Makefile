# Builds libmvit_b200.so (sm_100a only) and the oracle helpers.  `python -c "import __graft_entry__ as g; g.build()"`
# drives this; nvcc cross-compiles without a GPU.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -std=c++17 -O3 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v
SRC_DIR   := aicity_action_b200/csrc
OBJ_DIR   := build/obj
LIB       := aicity_action_b200/lib/libmvit_b200.so
SRCS      := $(wildcard $(SRC_DIR)/*.cu)
OBJS      := $(patsubst $(SRC_DIR)/%.cu,$(OBJ_DIR)/%.o,$(SRCS))
HDRS      := $(wildcard $(SRC_DIR)/*.cuh) include/mvit_b200.h

all: $(LIB)

$(OBJ_DIR)/%.o: $(SRC_DIR)/%.cu $(HDRS)
	@mkdir -p $(OBJ_DIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(OBJ_DIR)/$*.ptxas.log || (cat $(OBJ_DIR)/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	@mkdir -p $(dir $(LIB))
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS)

clean:
	rm -rf build $(LIB)

.PHONY: all clean
